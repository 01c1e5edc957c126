"""GPU side of the VTM acceptance harness (BASELINE configs[4]): predict PartitionMat files for synthetic sequences of the
VVC class shapes (D 416x240, C 832x480, B 1920x1080, A 3840x2160, plus two tiny multi-frame cases) with the B200 path
and store, under gpurun_out/vtm_cases/<name>/, the PartitionMat files and the float maps the decode consumed
(maps.npz: raw qt, bt, dire per component) -- run.py decodes the SAME maps with the unmodified reference
Map2Partition.py and feeds both file sets to the patched VTM-10.0 encoder.  The yuv files are not stored (gpurun_out is
size-limited): run.py regenerates them from the same seeds.

    python tools/vtm_acceptance/make_case.py [case ...]            (on the GPU box)
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
# name -> (width, height, frames, seed); None = the golden pipeline sequence of tests/cases.py
CASES = {"pipe_192x128": None, "mini_256x192": (256, 192, 2, 22), "classD_416x240": (416, 240, 1, 21),
         "classC_832x480": (832, 480, 1, 23), "classB_1920x1080": (1920, 1080, 1, 24), "classA_3840x2160": (3840, 2160, 1, 25)}


def case_frames(name):
    spec = CASES[name]
    if spec is None:
        y, u, v = cases.pipeline_frames()
        return cases.PIPE_W, cases.PIPE_H, cases.PIPE_F, y, u, v
    w, h, nf, seed = spec
    y, u, v = synth.synth_yuv420(w, h, nf, seed=seed)
    return w, h, nf, y, u, v


def main():
    names = sys.argv[1:] or list(CASES)
    out_root = os.path.join(ROOT, "gpurun_out", "vtm_cases")
    pp = PartitionPredictor(0, engine="tc")
    pp.load_pkls(os.path.join(ROOT, "trained_models"), qps=(QP,), missing_bd="seeded")
    summary = {}
    for name in names:
        w, h, nf, y, u, v = case_frames(name)
        d = os.path.join(out_root, name)
        os.makedirs(os.path.join(d, "PartitionMat"), exist_ok=True)
        res = pp.predict_frames(y, u, v, qps=(QP,), want_maps=True)
        sizes, maps = {}, {}
        for comp in COMPS:
            vals, qt, bt, dire, flags = res[(comp, QP)]
            path = pp.partition_path(os.path.join(d, "PartitionMat"), name, comp, QP)
            sizes[comp] = pp.write_partition_file(vals, path)
            maps[comp + "_qt"] = qt.cpu().numpy()
            maps[comp + "_bt"] = bt.cpu().numpy()
            maps[comp + "_dire"] = dire.cpu().numpy()
            maps[comp + "_flags"] = flags.cpu().numpy()
        np.savez_compressed(os.path.join(d, "maps.npz"), **maps)
        summary[name] = {"width": w, "height": h, "frames": nf, "qp": QP, "bytes": sizes,
                         "decode_report": pp.counts()["total"]}
    with open(os.path.join(out_root, "cases.json"), "w") as fp:
        json.dump(summary, fp, indent=1)
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
