"""BASELINE configs[3]: batch-size sweep over 64..16384 synthetic 128x128 CTUs (= 256..65536 64x64 blocks) per launch
chunk: luma + chroma Q+MSBD nets + post-process + decode at QP 32, inputs resident in HBM.  Prints one JSON line per size."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from pmp_vvc_tip2023_b200 import netspec, ops, synth  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [64, 256, 1024, 2048, 4096, 8192, 16384]
    pp = bench.load_predictor(0, "tc", 1200)
    by, bu, bv = synth.synth_blocks(4096, seed=11)
    lum = torch.from_numpy(by).unsqueeze(1)
    chr_ = torch.cat([torch.nn.functional.max_pool2d(lum.float(), 2).to(torch.uint8), torch.from_numpy(bu).unsqueeze(1),
                      torch.from_numpy(bv).unsqueeze(1)], 1)
    pk = bench.peaks()
    for ctus in sizes:
        nb = ctus * 4
        reps = (nb + 4095) // 4096
        lb = lum.repeat(reps, 1, 1, 1)[:nb].cuda().contiguous()
        cb = chr_.repeat(reps, 1, 1, 1)[:nb].cuda().contiguous()
        pp.chunk = min(nb, 8192)
        wl, wc = pp._wsets[("Luma", 32)], pp._wsets[("Chroma", 32)]

        def step():
            ops.run_component(wl[0], wl[1], True, lb, 1, 1, nb, pp.chunk, handle=pp.handle)
            ops.run_component(wc[0], wc[1], False, cb, 1, 1, nb, pp.chunk, handle=pp.handle)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        iters = max(2, min(20, 16384 // ctus))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        val = ctus / (ms * 1e-3)
        tf = val * netspec.FLOPS_PER_CTU / 1e12
        print(json.dumps({"ctus_per_launch": ctus, "blocks": nb, "chunk": pp.chunk, "ms": ms, "ctu_per_s": val,
                          "algorithmic_tflops": tf, "frac_of_sustained_bf16_peak_issued_x3": 3 * tf / pk["bf16_sustained"]}), flush=True)
        del lb, cb


if __name__ == "__main__":
    main()
