"""GPU side of the VTM acceptance harness (BASELINE configs[4]): predict PartitionMat files for small synthetic
sequences of several class shapes with the B200 path and store yuv + files under gpurun_out/vtm_cases/<name>/.

    python tools/vtm_acceptance/make_case.py            (on the GPU box)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from pmp_vvc_tip2023_b200 import synth  # noqa: E402
from pmp_vvc_tip2023_b200.pipeline import COMPS, PartitionPredictor  # noqa: E402
from tests import cases  # noqa: E402

QP = 32
CASES = {"pipe_192x128": None, "classD_416x240": (416, 240, 1, 21), "mini_256x192": (256, 192, 2, 22)}


def main():
    out_root = os.path.join(ROOT, "gpurun_out", "vtm_cases")
    pp = PartitionPredictor(0, engine="tc")
    pp.load_pkls(os.path.join(ROOT, "trained_models"), qps=(QP,), missing_bd="seeded")
    summary = {}
    for name, spec in CASES.items():
        if spec is None:
            w, h, nf = cases.PIPE_W, cases.PIPE_H, cases.PIPE_F
            y, u, v = cases.pipeline_frames()
        else:
            w, h, nf, seed = spec
            y, u, v = synth.synth_yuv420(w, h, nf, seed=seed)
        d = os.path.join(out_root, name)
        os.makedirs(os.path.join(d, "PartitionMat"), exist_ok=True)
        with open(os.path.join(d, name + ".yuv"), "wb") as fp:
            for f in range(nf):
                fp.write(y[f].tobytes()); fp.write(u[f].tobytes()); fp.write(v[f].tobytes())
        res = pp.predict_frames(y, u, v, qps=(QP,))
        sizes = {}
        for comp in COMPS:
            path = pp.partition_path(os.path.join(d, "PartitionMat"), name, comp, QP)
            sizes[comp] = pp.write_partition_file(res[(comp, QP)], path)
        summary[name] = {"width": w, "height": h, "frames": nf, "qp": QP, "bytes": sizes}
    with open(os.path.join(out_root, "cases.json"), "w") as fp:
        json.dump(summary, fp, indent=1)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
