"""Drop-in for the reference driver ``Inference_QBD.py``: same function names and CLI flags
(``--jobID --inputDir --outDir --batchSize --startSeqID --seqNum``, Inference_QBD.py:257-267), same output tree
``<outDir>/<jobID>/PartitionMat/<seq>_<comp>_QP<qp>_PartitionMat.txt`` and ``Time_Sta_*.txt``.

The reference hard-codes ``Training_Sequences.txt``, ``.\\per-sequence`` and ``./CTU_Models/`` (:50,:162,:219);
here they are the defaults of extra flags (``--seqInfo --cfgDir --modelDir --ssRatio --gpus``).  Frames are cut,
predicted, decoded and formatted on the GPU; with ``--gpus N`` the frames of a sequence are split into N contiguous
ranges, one worker thread + handle per GPU, and the per-GPU text segments are concatenated in frame order.
"""
import argparse
import os
import time

import numpy as np
import torch

from .Metrics import inference_pre_QBD, seq_post_process       # noqa: F401  (reference re-exports)
from .pipeline import COMPS, PartitionPredictor
from .weights import load_pretrain_model, remove_prefix          # noqa: F401

SAVE_MID_RESULT = False
POST_PROCESS = True
SSRatio = 30


def load_sequences_info(seqs_info_path="Training_Sequences.txt", ss_ratio=None):
    """Inference_QBD.py:48-76 on any CSV in the VVC_Test_Sequences.txt format (terminated by '#end!!!!')."""
    ss = SSRatio if ss_ratio is None else ss_ratio
    rows = []
    with open(seqs_info_path, "r") as fp:
        for line in fp:
            if "end!!!!" in line:
                break
            if line.strip():
                rows.append(line.rstrip("\n").split(","))
    data = np.array(rows)
    names, paths = data[:, 0], data[:, 1]
    w, h, nf = data[:, 2].astype(np.int64), data[:, 3].astype(np.int64), data[:, 4].astype(np.int64)
    sub = [(int(n) + ss - 1) // ss for n in nf]
    blocks = [int((w[i] // 64) * (h[i] // 64) * sub[i]) for i in range(len(sub))]
    return names, paths, w, h, nf, sub, blocks


def import_yuv420(file_path, width, height, frm_num, SubSampleRatio=1, is10bit=False):
    """Inference_QBD.py:78-102: planar 4:2:0, 8-bit or 16-bit LE samples, temporal subsampling by seek."""
    pix = width * height
    dt = np.uint16 if is10bit else np.uint8
    idx = list(range(0, frm_num, SubSampleRatio))
    y = np.zeros((len(idx), height, width), dt)
    u = np.zeros((len(idx), height // 2, width // 2), dt)
    v = np.zeros((len(idx), height // 2, width // 2), dt)
    with open(file_path, "rb") as fp:
        for k, i in enumerate(idx):
            fp.seek(i * pix * 3 if is10bit else i * pix * 3 // 2, 0)
            y[k] = np.fromfile(fp, dtype=dt, count=pix).reshape(height, width)
            u[k] = np.fromfile(fp, dtype=dt, count=pix // 4).reshape(height // 2, width // 2)
            v[k] = np.fromfile(fp, dtype=dt, count=pix // 4).reshape(height // 2, width // 2)
    return y, u, v


def output_block_yuv(file_path, width, height, block_size, in_overlap, numfrm, SubSampleRatio, is10bit=False,
                     save_path=None):
    """Inference_QBD.py:104-149 (host arrays out, as in the reference; the block cutting itself runs on the GPU)."""
    if block_size != 64 or in_overlap != 4:
        raise NotImplementedError("the nets are defined for block_size=64, in_overlap=4 (Inference_QBD.py:190)")
    y, u, v = import_yuv420(file_path, width, height, numfrm, SubSampleRatio, is10bit=is10bit)
    from . import ops
    dev = torch.device("cuda", torch.cuda.current_device())
    lb, cb = ops.cut_blocks(*(torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).to(dev) for a in (y, u, v)))
    block_y = lb[:, 0].cpu().numpy()
    block_u, block_v = cb[:, 1].cpu().numpy(), cb[:, 2].cpu().numpy()
    if save_path is not None:
        with open(save_path, "wb") as fp:
            for i in range(block_y.shape[0]):
                fp.write(block_y[i].tobytes()); fp.write(block_u[i].tobytes()); fp.write(block_v[i].tobytes())
    return block_y, block_u, block_v


def _parse_cfg(path):
    seq_path, is10bit = None, False
    with open(path) as fp:
        for line in fp:
            body = line.rstrip("\n").split("#")[0].replace(" ", "")
            if "InputFile" in body:
                seq_path = body.split(":", 1)[1]
            elif "InputBitDepth" in body:
                is10bit = body.split(":", 1)[1] == "10"
    return seq_path, is10bit


def _predict_shard(pred, y, u, v, qps, want_raw=False):
    """One GPU's frame range: cut once, then per (component, QP) the nets + post-process + decode + frame assembly
    (timed with CUDA events, the counterpart of the reference's per-(QP, comp) inference clock, Inference_QBD.py:210-227)
    and the text formatting + device->host copy (wall clock; the reference's post-process clock, :229-241).
    Returns ({(comp, qp): uint8 array}, {(comp, qp): (net_s, post_s)}, counts, {(comp, qp): int8 array} when want_raw);
    the arrays are views of the predictor's pinned staging buffers (valid until its next call)."""
    from . import ops
    texts, times, raw = {}, {}, {}
    f, hgt, wid = y.shape
    bh, bw = hgt // 64, wid // 64
    dev = pred.device
    with torch.cuda.device(dev):
        lb, cb = pred.cut(y, u, v)
        for comp in COMPS:
            blocks = lb if comp == "Luma" else cb
            for qp in qps:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                vals = pred.predict_blocks(comp, qp, blocks, f, bh, bw)
                e1.record()
                e1.synchronize()
                t0 = time.time()
                # pinned staging buffers, reused across sequences; the writer consumes the views before the next sequence
                texts[(comp, qp)] = pred.to_host_pinned(("text", comp, qp), ops.format_text(vals, handle=pred.handle)).numpy()
                if want_raw:
                    raw[(comp, qp)] = pred.to_host_pinned(("raw", comp, qp), vals).numpy()
                times[(comp, qp)] = (e0.elapsed_time(e1) * 1e-3, time.time() - t0)
        counts = pred.counts()
    return texts, times, counts, raw


def _predict_pieces(pred, y, u, v, qps, calls, want_raw=False):
    """One GPU's shard = calls [([qp_index...], frame_lo, frame_hi)] (sharding.group_calls); a GPU holds at most one frame
    range per QP, so the per-(comp, qp) results of the calls merge without collisions."""
    texts, times, raw = {}, {}, {}
    counts = None
    pred.last_flags = {}            # the decode report below covers exactly this shard's (comp, qp) results
    for qis, a, b in calls:
        t, tm, counts, r = _predict_shard(pred, y[a:b], u[a:b], v[a:b], tuple(qps[i] for i in qis), want_raw)
        texts.update(t); times.update(tm); raw.update(r)        # counts: last_flags accumulates over the calls
    return texts, times, counts, raw


@torch.no_grad()
def inference_VVC_seqs(args):
    from concurrent.futures import ThreadPoolExecutor
    save_dir = os.path.join(args.outDir, args.jobID, "PartitionMat")
    os.makedirs(save_dir, exist_ok=True)
    ss = getattr(args, "ssRatio", SSRatio)
    gpus = max(1, min(getattr(args, "gpus", 1), torch.cuda.device_count()))
    model_dir = getattr(args, "modelDir", "./CTU_Models/")
    missing_bd = getattr(args, "missingBD", "error")
    qps = (22, 27, 32, 37)
    names, paths, widths, heights, frmnums, subs, _ = load_sequences_info(getattr(args, "seqInfo", "Training_Sequences.txt"), ss)
    preds = [PartitionPredictor(g, engine=getattr(args, "engine", "tc"), chunk=max(args.batchSize, 256),
                                near_tol=getattr(args, "nearTol", 1e-2), tc_dtype=getattr(args, "tcDtype", "fp16"))
             for g in range(gpus)]
    for p in preds:
        p.load_pkls(model_dir, missing_bd=missing_bd)
    nseq = args.seqNum
    t_block, t_net, t_post = np.zeros(nseq), np.zeros((nseq, 4, 2)), np.zeros((nseq, 4, 2))
    for seq_id in range(args.startSeqID, args.startSeqID + nseq):
        s = seq_id - args.startSeqID
        name, width, height = names[seq_id], int(widths[seq_id]), int(heights[seq_id])
        seq_path_name = paths[seq_id][:-4] if paths[seq_id].endswith(".yuv") else paths[seq_id]
        cfg = os.path.join(getattr(args, "cfgDir", r".\per-sequence"), name + ".cfg")
        if os.path.exists(cfg):
            seq_path, is10bit = _parse_cfg(cfg)
        else:
            seq_path, is10bit = os.path.join(args.inputDir, paths[seq_id]), "10bit" in paths[seq_id]
        print(name)
        t0 = time.time()
        y, u, v = import_yuv420(seq_path, width, height, int(frmnums[seq_id]), ss, is10bit=is10bit)
        nf = y.shape[0]
        t_block[s] = time.time() - t0
        # shards: (QP, frame) pairs split evenly over the GPUs (sharding.py); --shard frames = whole frame ranges, all QPs
        from .sharding import frame_shards, group_calls, qp_frame_shards
        shard_fn = frame_shards if getattr(args, "shard", "frame-qp") == "frames" else qp_frame_shards
        pieces = {g: shard_fn(nf, len(qps), gpus, g) for g in range(gpus)}
        shards = [g for g in range(gpus) if pieces[g]]
        # one worker thread per GPU; .result() re-raises whatever a worker raised (OOM, PmpError, CUDA error): a failed
        # shard must never turn into a silently truncated PartitionMat file
        with ThreadPoolExecutor(max_workers=max(1, len(shards))) as pool:
            futs = {g: pool.submit(_predict_pieces, preds[g], y, u, v, qps, group_calls(pieces[g]),
                                   getattr(args, "binaryOut", False)) for g in shards}
            results = {g: f.result() for g, f in futs.items()}
        # every file must be covered exactly once, in frame order, by the rank-ordered segments of its QP
        for qi in range(len(qps)):
            segs = [(a, b) for g in shards for q2, a, b in pieces[g] if q2 == qi]
            if not segs or segs[0][0] != 0 or segs[-1][1] != nf or any(segs[i][1] != segs[i + 1][0] for i in range(len(segs) - 1)):
                raise RuntimeError("shards of QP %d do not tile the %d frames: %s" % (qps[qi], nf, segs))
        tot = {"blocks": 0, "near_tie_blocks": 0, "near_threshold_blocks": 0, "near_threshold_blocks_tight": 0,
               "fp16_saturation_events": 0}
        for g in shards:
            for key in tot:
                tot[key] += results[g][2]["total"][key]
        near_tol = getattr(args, "nearTol", 1e-2)
        print("Decode report: %d block-QPs, %d float32 near-tie argmins, %d with a map value within %g of a decision "
              "threshold (%d within %g), %d fp16 saturation events" % (
                  tot["blocks"], tot["near_tie_blocks"], tot["near_threshold_blocks"], near_tol,
                  tot["near_threshold_blocks_tight"], 0.01 * near_tol, tot["fp16_saturation_events"]))
        if tot["fp16_saturation_events"]:
            raise RuntimeError("activations left the fp16 range (%d events): results are clamped; rerun with --tcDtype bf16"
                               % tot["fp16_saturation_events"])
        for ci, comp in enumerate(COMPS):
            for qi, qp in enumerate(qps):
                path = PartitionPredictor.partition_path(save_dir, seq_path_name, comp, qp)
                print("Save:", path)
                t0 = time.time()
                with open(path, "wb") as fp:
                    for g in shards:
                        if any(q2 == qi for q2, _, _ in pieces[g]):
                            fp.write(results[g][0][(comp, qp)])       # KeyError if a shard did not deliver
                if getattr(args, "binaryOut", False):
                    # opt-in raw int8 form for the VTM-side binary reader (tools/vtm_reader/pmp_partition_reader.h)
                    from .partition_io import bin_header, values_per_frame
                    r_, c_, _ = values_per_frame(y.shape[1], y.shape[2])
                    with open(path[:-4] + ".bin", "wb") as fp:
                        fp.write(bin_header(y.shape[0], r_, c_))
                        for g in shards:
                            if any(q2 == qi for q2, _, _ in pieces[g]):
                                fp.write(results[g][3][(comp, qp)])
                # shards run concurrently: the sequence's time is the slowest GPU's
                owners = [g for g in shards if (comp, qp) in results[g][1]]
                t_net[s, qi, ci] = max(results[g][1][(comp, qp)][0] for g in owners)
                t_post[s, qi, ci] = max(results[g][1][(comp, qp)][1] for g in owners) + (time.time() - t0)
    log = os.path.join(args.outDir, args.jobID, "Time_Sta_%d_%d.txt" % (args.startSeqID, args.startSeqID + nseq))
    with open(log, "w") as fp:
        for s in range(nseq):
            for q in range(4):
                fp.write(",".join(str(x) for x in (t_block[s], t_net[s, q, 0], t_net[s, q, 1], t_post[s, q, 0],
                                                   t_post[s, q, 1])) + ",\n")
    print("Sum time:", np.sum(t_block) + np.sum(t_net) + np.sum(t_post))
    for p in preds:
        p.close()


def build_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('--jobID', type=str, default='0000')
    parser.add_argument('--inputDir', type=str, default='/input/')
    parser.add_argument('--outDir', type=str, default='/output/')
    parser.add_argument('--batchSize', default=200, type=int, help='batch size')
    parser.add_argument('--startSeqID', default=0, type=int, help='QP start ID')
    parser.add_argument('--seqNum', default=22, type=int, help='test QP number')
    # extras: the reference hard-codes these
    parser.add_argument('--seqInfo', type=str, default='Training_Sequences.txt')
    parser.add_argument('--cfgDir', type=str, default=r'.\per-sequence')
    parser.add_argument('--modelDir', type=str, default='./CTU_Models/')
    parser.add_argument('--ssRatio', type=int, default=SSRatio)
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--engine', type=str, default='tc', choices=['tc', 'simt'])
    parser.add_argument('--missingBD', type=str, default='error', choices=['error', 'seeded'])
    parser.add_argument('--tcDtype', type=str, default='fp16', choices=['fp16', 'bf16'],
                        help='16-bit operand format of the split-precision tensor-core engine')
    parser.add_argument('--nearTol', type=float, default=1e-2, help='tolerance of the near-threshold block count')
    parser.add_argument('--shard', type=str, default='frame-qp', choices=['frame-qp', 'frames'],
                        help='multi-GPU work split: (QP, frame) pairs evenly (default) or whole frame ranges with all QPs')
    parser.add_argument('--binaryOut', action='store_true',
                        help='also write <seq>_<comp>_QP<qp>_PartitionMat.bin (raw int8) for tools/vtm_reader')
    return parser


if __name__ == '__main__':
    a = build_parser().parse_args()
    t = time.time()
    inference_VVC_seqs(a)
    print('Total inference time:', time.time() - t)
