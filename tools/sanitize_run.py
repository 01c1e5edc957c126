"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one 192x128 frame, both engines."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pmp_vvc_tip2023_b200 import synth, ops
from pmp_vvc_tip2023_b200.pipeline import PartitionPredictor
y, u, v = synth.synth_yuv420(192, 128, 1, seed=7)
for engine in (sys.argv[1:] or ["simt", "tc"]):
    pp = PartitionPredictor(0, engine=engine, chunk=4)
    pp.load_seeded(qps=(32,))
    res = pp.predict_frames(y, u, v, qps=(32,))
    for k, t in res.items():
        txt = ops.format_text(t)
        print(engine, k, t.shape, int(txt.numel()))
torch.cuda.synchronize()
print("done")
