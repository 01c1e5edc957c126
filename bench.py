#!/usr/bin/env python
"""Benchmark of the partition-map prediction hot path (BASELINE.json metric: CTUs/sec & 4K frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--engine tc|simt]

Workload (BASELINE.json configs[1]): a synthetic 1920x1080 10-bit 4:2:0 10-frame sequence (480 64x64 blocks =
120 CTUs per frame), luma + chroma QT+MTT nets at QP 22/27/32/37, QT post-process, map-to-partition decode and frame
assembly.  One "step" = one pass of the whole path over that sequence: 10 frames x 120 CTUs x 4 QPs = 4800 CTU-QP
units (a unit = one 128x128 CTU through luma Q+MSBD, chroma Q+MSBD and both decodes for one QP = 37.164 GFLOP).
N > 1: every rank runs its own sequence (frame sharding, no data-path collective) -> weak scaling.

`value`  : units/s with the frames already resident in HBM (device timing, CUDA events, max over ranks).
`e2e`    : the same through the public API (PartitionPredictor.predict_frames) from pinned HOST frames, H2D of the
           frames and D2H of the int8 partition vectors inside the timed region.
`roofline`: the conv kernel class that dominates the step, algorithmic FLOPs / CUDA-event time of its launches over a
           second pass of the same K steps (per-launch event pairs; `value` is timed without them), against the
           measured bf16 peak in MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port of the reference's CPU path (PyTorch CPU fp32 nets + NumPy post-process/decode),
           timed on this host on a bounded sample of the same workload (rank 0, N=1 only).
`--impl reference`: times that CPU path alone (same metric/config), see reference_main().
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WIDTH, HEIGHT, FRAMES = 1920, 1080, 10
QPS = (22, 27, 32, 37)
COMPS = ("Luma", "Chroma")
METRIC = "CTUs/sec (128x128 CTU, luma+chroma QT+MTT nets + post-process + Map2Partition decode, per QP)"
UNIT = "CTU/s"


def workload_config(engine):
    bh, bw = HEIGHT // 64, WIDTH // 64
    return {"workload": "synthetic 1920x1080 10-bit 4:2:0, %d frames, Luma+Chroma x QP 22/27/32/37 (BASELINE configs[1])" % FRAMES,
            "blocks_per_frame": bh * bw, "ctus_per_frame": bh * bw // 4, "frames": FRAMES, "qps": list(QPS),
            "units_per_step": FRAMES * bh * bw * len(QPS) // 4, "engine": engine,
            "l2": "no flush: per-step working set (activation arena, GBs) >> 126 MB L2",
            "weights": "Q nets: reference trained .pkl; MSBD nets: seeded random (trained *_BD_*.pkl absent offline)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.proc.wait()
        sm, mx, reasons, pw = [], [], set(), []
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def load_predictor(dev, engine, chunk):
    from pmp_vvc_tip2023_b200.pipeline import PartitionPredictor
    pp = PartitionPredictor(dev, engine=engine, chunk=chunk)
    pp.load_pkls(os.path.join(ROOT, "trained_models"), missing_bd="seeded")
    return pp


def make_frames(seed):
    from pmp_vvc_tip2023_b200 import synth
    # one synthetic frame generator call per distinct frame is slow in numpy at 1080p: build 2 and alternate with shifts
    y, u, v = synth.synth_yuv420(WIDTH, HEIGHT, 2, seed=seed)
    idx = [i % 2 for i in range(FRAMES)]
    ys = np.stack([np.roll(y[k], 8 * i, axis=1) for i, k in enumerate(idx)])
    us = np.stack([np.roll(u[k], 4 * i, axis=1) for i, k in enumerate(idx)])
    vs = np.stack([np.roll(v[k], 4 * i, axis=1) for i, k in enumerate(idx)])
    return ys, us, vs


# ------------------------------------------------------------------------------------------------------
# CPU path (oracle port of the reference): used for cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------------
class CpuPath:
    def __init__(self):
        from oracle import decode_ref, nets_ref, postproc_ref
        from pmp_vvc_tip2023_b200 import synth
        from pmp_vvc_tip2023_b200.weights import load_reference_pkl
        self.nets_ref, self.postproc_ref, self.decode_ref = nets_ref, postproc_ref, decode_ref
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = {}
        for comp in COMPS:
            sdq = load_reference_pkl(os.path.join(ROOT, "trained_models", "%s_Q_32.pkl" % comp))
            sdb = {k: torch.from_numpy(v) for k, v in synth.seeded_state_dict(comp + "_MSBD", 1032 + (0 if comp == "Luma" else 500)).items()}
            self.sd[comp] = (sdq, sdb)
        y, u, v = make_frames(0)
        self.blocks = nets_ref.cut_blocks(y[:1], u[:1], v[:1], True)      # 480 blocks of frame 0

    def run(self, nblocks, lo=0):
        """nblocks blocks (luma + chroma, one QP) through nets -> post-process -> NumPy decode.  Returns (s_nets, s_post)."""
        by, bu, bv = (b[lo:lo + nblocks] for b in self.blocks)
        t_net = t_post = 0.0
        for comp in COMPS:
            luma = comp == "Luma"
            t0 = time.perf_counter()
            x = torch.from_numpy(by.astype(np.float32)).unsqueeze(1) if luma else self.nets_ref.chroma_net_input(by, bu, bv)
            qt, bt, dire = self.nets_ref.predict_maps(self.sd[comp][0], self.sd[comp][1], x, luma, batch=200)
            t1 = time.perf_counter()
            qi = self.postproc_ref.eli_structural_error(qt.numpy())[:, 0]
            btn, din = bt.numpy(), dire.numpy()
            for b in range(qi.shape[0]):       # single Python thread, as Map2Partition.py:389-399 runs it
                self.decode_ref.map_to_partition(qi[b], btn[b], din[b], 1 if luma else 2)
            t2 = time.perf_counter()
            t_net += t1 - t0
            t_post += t2 - t1
        return t_net, t_post


def cpu_baseline_sample(budget_s=20.0):
    cp = CpuPath()
    tn, tp = cp.run(8)                                     # calibrate
    per_block = (tn + tp) / 8
    n = int(max(8, min(480, budget_s / max(per_block, 1e-6))))
    tn, tp = cp.run(n)
    return {"value": (n / 4.0) / (tn + tp), "unit": UNIT, "cores": cp.cores, "kind": "port",
            "sample": "%d blocks (=%g CTUs) of frame 0, luma+chroma, QP 32: torch-CPU fp32 nets (%d threads, batch 200) "
                      "%.2f s + NumPy post-process/decode (1 thread, as the reference runs it) %.2f s" % (n, n / 4.0, cp.cores, tn, tp),
            "nets_s": tn, "post_s": tp}


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cp = CpuPath()
    tn, tp = cp.run(8)
    per_block = (tn + tp) / 8
    total_budget = 150.0
    n = int(max(4, min(480, total_budget / max(args.steps + args.warmup, 1) / max(per_block, 1e-6))))
    for _ in range(args.warmup):
        cp.run(n)
    t0 = time.perf_counter()
    tn = tp = 0.0
    for _ in range(args.steps):
        a, b = cp.run(n)
        tn += a; tp += b
    el = time.perf_counter() - t0
    value = args.steps * (n / 4.0) / el
    sample = ("%d blocks (=%g CTUs) per step of frame 0, luma+chroma, QP 32; nets %.2f s (torch CPU fp32, %d threads), "
              "post-process+decode %.2f s (NumPy, 1 thread)" % (n, n / 4.0, tn, cp.cores, tp))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * el / max(args.steps, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config("reference-cpu"),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cp.cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "reference is Python/PyTorch (not installable as a binary): its CPU path is the oracle port, "
                   "pinned to the reference's own outputs by tests/golden"}
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--engine", type=str, default="tc", choices=["tc", "simt"])
    ap.add_argument("--chunk", type=int, default=2400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_main(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from pmp_vvc_tip2023_b200 import netspec
    pp = load_predictor(local, args.engine, args.chunk)
    h = pp.handle
    y, u, v = make_frames(100 + rank)
    hy, hu, hv = (torch.from_numpy(a.view(np.int16)).pin_memory() for a in (y, u, v))
    dy, du, dv = (t.cuda() for t in (hy, hu, hv))
    bh, bw = HEIGHT // 64, WIDTH // 64
    units = FRAMES * bh * bw * len(QPS) / 4.0
    per = 2 * (bh * 16) * (bw * 16) + (bh * 8) * (bw * 8) + 3 * (bh * 16) * (bw * 16)
    host_out = {(c, q): torch.empty((FRAMES, per), dtype=torch.int8).pin_memory() for c in COMPS for q in QPS}

    def step_device():
        return pp.predict_frames(dy, du, dv, qps=QPS)

    def step_e2e():
        pp.predict_frames(hy, hu, hv, qps=QPS, host_out=host_out)      # D2H of each component overlaps the next one
        pp.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = h.launch_count()
    ms = timed(step_device, args.steps)                 # `value`: no per-launch events in the stream
    launches = h.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # roofline pass: the same K steps again with a CUDA-event pair around every launch (the events sit between the
    # kernels, which switches off the programmatic-dependent-launch overlap, hence a separate pass)
    prof, ms_prof = {}, None
    if not args.no_profile:
        h.profile(2)
        ms_prof = timed(step_device, args.steps)
        prof = h.profile_read()
        h.profile(0)

    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    value = world * units * args.steps / (ms * 1e-3)
    e2e_value = world * units * args.steps / (ms_e2e * 1e-3)
    dom = "conv_tc" if args.engine == "tc" and prof.get("conv_tc", {}).get("ms", 0) > 0 else "conv_simt"
    roof = None
    if not args.no_profile and prof.get(dom, {}).get("ms", 0) > 0:
        p = prof[dom]
        ach = p["flops"] / (p["ms"] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r01b_traffic.json")
        if dom == "conv_tc" and os.path.exists(tpath):
            # dram__bytes_read+write per conv_tc launch from an ncu pass over 480-block launches (committed summary),
            # scaled to this run's blocks per launch: HBM bytes per launch of the dominant kernel class
            tj = json.load(open(tpath))
            traffic = tj["conv_tc_dram_bytes_per_launch_per_block"] * min(args.chunk, FRAMES * (HEIGHT // 64) * (WIDTH // 64))
            traffic_src = tj["source"] + "; scaled from 480 to %d blocks per launch" % min(args.chunk, FRAMES * (HEIGHT // 64) * (WIDTH // 64))
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_sustained"], "traffic": traffic, "traffic_src": traffic_src,
                "peak_src": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches": p["launches"], "avg_launch_ms": p["ms"] / max(p["launches"], 1),
                "algorithmic_flops_per_launch": p["flops"] / max(p["launches"], 1),
                "share_of_step": p["ms"] / ms_prof, "profiled_pass_ms_per_step": ms_prof / args.steps,
                "note": "algorithmic FLOPs = 2*MACs of the fp32 reference convs; the TC engine issues 3 fp16 MMA passes per "
                        "MAC (split precision), so issued tensor work = 3x this"}
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f16x3(split)+f32acc" if args.engine == "tc" else "f32", "data": "synthetic",
           "config": workload_config(args.engine),
           "frames_per_s_1080p_all_qps": world * FRAMES * args.steps / (ms * 1e-3),
           "frame_qp_per_s_4k_equiv": value / 495.0,
           "algorithmic_tflops": value * netspec.FLOPS_PER_CTU / 1e12,
           "gpu_launches": int(launches), "clocks": clocks,
           "e2e": {"value": e2e_value, "unit": UNIT,
                   "h2d_bytes_per_step": int(hy.numel() * 2 + hu.numel() * 2 + hv.numel() * 2),
                   "d2h_bytes_per_step": int(sum(t.numel() for t in host_out.values())),
                   "ms_per_step": ms_e2e / args.steps},
           "roofline": roof,
           "kernel_classes_ms_per_step": {k: p["ms"] / args.steps for k, p in prof.items() if p["launches"]},
           # the integer / element-wise kernel classes against the HBM roofline: algorithmic bytes
           # over the CUDA-event time of their launches in the profiled pass
           "hbm_kernels": {k: {"achieved_gbs": p["bytes"] / (p["ms"] * 1e-3) / 1e9, "peak_gbs": pk["hbm_gbs"],
                               "frac": p["bytes"] / (p["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], "launches": p["launches"]}
                           for k, p in prof.items() if p["launches"] and p["bytes"] > 0 and p["ms"] > 0 and not k.startswith("conv")}}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample()
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
