"""Short single-GPU workload for ncu: one 1080p frame (480 blocks), luma + chroma, QP 32, TC engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    pp = bench.load_predictor(0, "tc", int(sys.argv[2]) if len(sys.argv) > 2 else 480)
    y, u, v = bench.make_frames(100)
    for _ in range(2):
        pp.predict_frames(y[:frames], u[:frames], v[:frames], qps=(32,))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
