"""Deterministic test-case generators shared by oracle/gen_golden.py (which freezes the
reference's answers) and the tests (which regenerate the same inputs from the seeds)."""
import numpy as np

from pmp_vvc_tip2023_b200 import synth

PIPE_W, PIPE_H, PIPE_F = 192, 128, 2


def msbd_seed(comp, qp):
    return 1000 + qp + (0 if comp == "Luma" else 500)


def _consistent_qt(rng, n):
    """Random qt maps of the shape eli_structual_error emits (4x4 ints upsampled x2)."""
    from oracle import postproc_ref
    raw = (rng.standard_normal((n, 1, 8, 8)) * 1.1 + 1.2).astype(np.float32)
    # make them blocky so that all qt depths occur
    raw = np.repeat(np.repeat(raw[:, :, ::4, ::4], 4, 2), 4, 3) * 0.7 + raw * 0.3
    return postproc_ref.eli_structural_error(raw)[:, 0]


def _smooth_field(rng, n, lo, hi, cells):
    g = rng.uniform(lo, hi, size=(n, 3, cells, cells))
    reps = 16 // cells
    f = np.repeat(np.repeat(g, reps, 2), reps, 3)
    f += 0.25 * rng.standard_normal(f.shape)
    return f.astype(np.float32)


def decode_cases():
    """name -> (qt [n,8,8] f32, bt [n,3,16,16] f32, dire [n,3,16,16] f32, chroma_factor)."""
    out = {}
    for chroma in (False, True):
        tag = "chroma" if chroma else "luma"
        cf = 2 if chroma else 1
        for sig in (0.0, 0.15, 0.3):
            qt, bt, dire = synth.structured_maps(96, seed=21 + int(sig * 100), sigma=sig, chroma=chroma)
            out["struct_%s_s%02d" % (tag, int(sig * 100))] = (qt, bt, dire, cf)
        # quantised noise: exact ties and x.5 rounding cases
        qt, bt, dire = synth.structured_maps(96, seed=77, sigma=0.35, chroma=chroma)
        out["quant_%s" % tag] = (qt, np.round(bt * 4) / 4, np.round(dire * 4) / 4, cf)
        # unstructured smooth fields on consistent qt maps
        rng = np.random.default_rng([5, int(chroma)])
        n = 64
        qt = _consistent_qt(rng, n)
        b0 = _smooth_field(rng, n, 0.0, 2.2, 4)[:, 0]
        b1 = b0 + np.abs(_smooth_field(rng, n, 0.0, 1.2, 8)[:, 0])
        b2 = b1 + np.abs(_smooth_field(rng, n, 0.0, 1.2, 8)[:, 0])
        bt = np.stack([b0, b1, b2], 1).astype(np.float32)
        dire = _smooth_field(rng, n, -1.3, 1.3, 4)
        out["smooth_%s" % tag] = (qt, bt, dire, cf)
        # arbitrary (possibly inconsistent) qt ints, wide-range maps, degenerate blocks
        rng = np.random.default_rng([6, int(chroma)])
        n = 48
        qt = rng.integers(0, 4, size=(n, 8, 8)).astype(np.float32)
        qt[:8] = np.repeat(np.repeat(rng.integers(0, 4, size=(8, 2, 2)), 4, 1), 4, 2)
        bt = (rng.standard_normal((n, 3, 16, 16)) * 2.0 + 1.0).astype(np.float32)
        dire = (rng.standard_normal((n, 3, 16, 16)) * 1.5).astype(np.float32)
        qt[0], bt[0], dire[0] = 0, 0, 0
        qt[1], bt[1], dire[1] = 3, 5.0, 1.0
        qt[2], bt[2], dire[2] = 0, -3.0, -1.0
        qt[3], bt[3], dire[3] = 1, 1.0, 0.5
        qt[4], bt[4], dire[4] = 2, 2.5, -0.5
        out["wild_%s" % tag] = (qt, bt, dire, cf)
    return out


def postproc_inputs():
    rng = np.random.default_rng(99)
    a = (rng.standard_normal((700, 1, 8, 8)) * 1.2 + 1.3)
    b = np.repeat(np.repeat(rng.uniform(-0.6, 3.8, size=(500, 1, 4, 4)), 2, 2), 2, 3) + \
        0.2 * rng.standard_normal((500, 1, 8, 8))
    c = np.round(rng.uniform(-1, 4.5, size=(300, 1, 8, 8)) * 2) / 2          # exact .5 ties
    d = np.zeros((8, 1, 8, 8))
    d[1] = 0.49; d[2] = 0.5; d[3] = 1.5; d[4] = 2.5; d[5] = 3.7; d[6] = -2; d[7, 0, :4] = 2.2
    # nearly-all-zero maps (13..15 zeros after pooling)
    e = np.zeros((64, 1, 8, 8))
    for i in range(64):
        k = rng.integers(1, 5)
        idx = rng.integers(0, 8, size=(k, 2))
        e[i, 0, idx[:, 0], idx[:, 1]] = rng.uniform(0.6, 3.4, size=k)
    return np.concatenate([a, b, c, d, e]).astype(np.float32)


def net_blocks():
    """8 luma+chroma blocks (uint8) cut from one synthetic 416x240 10-bit frame."""
    from oracle import nets_ref
    y, u, v = synth.synth_yuv420(416, 240, 1, seed=1)
    by, bu, bv = nets_ref.cut_blocks(y, u, v, True)
    idx = [0, 2, 5, 6, 9, 11, 14, 17]
    return by[idx], bu[idx], bv[idx]


def pipeline_frames():
    return synth.synth_yuv420(PIPE_W, PIPE_H, PIPE_F, seed=4)


def value_block_ids(frames, bh, bw):
    """Block id (frame-major raster) of every value of the per-frame PartitionMat vectors (hor | ver | qt | dire,
    Map2Partition.py:401-412), flattened over frames: maps a differing file line back to its 64x64 block."""
    R, C = 16 * bh, 16 * bw
    r16 = (np.arange(R)[:, None] // 16) * bw + (np.arange(C)[None, :] // 16)
    r8 = (np.arange(R // 2)[:, None] // 8) * bw + (np.arange(C // 2)[None, :] // 8)
    per = np.concatenate([r16.reshape(-1), r16.reshape(-1), r8.reshape(-1), np.tile(r16.reshape(-1), 3)])
    return (np.arange(frames)[:, None] * (bh * bw) + per[None, :]).reshape(-1)


def threshold_distance(qt, bt, dire):
    """Per block: the smallest distance of any map value to a decision threshold of the integer path -- the 2x2-pooled
    raw qt to 0.5/1.5/2.5 (Metrics.py:631-632), bt to k+0.5 (np.round, Map2Partition.py:104), dire to +-0.5 (:105)."""
    n = bt.shape[0]
    q = qt.reshape(n, 4, 2, 4, 2).max(axis=(2, 4)).reshape(n, -1)
    dq = np.where((q > 0) & (q < 3), np.abs(q - np.floor(q) - 0.5), np.inf).min(axis=1)
    b = bt.reshape(n, -1)
    db = np.abs(b - np.floor(b) - 0.5).min(axis=1)
    dd = np.abs(np.abs(dire.reshape(n, -1)) - 0.5).min(axis=1)
    return np.minimum(np.minimum(dq, db), dd)


# non-default Map_to_Partition constructor thresholds (Map2Partition.py:100) exercised against the reference
LAMB_SETS = {"a": (0.6, 0.8, 1.2, 0.4, 0.6), "b": (0.9, 0.5, 2.0, 0.2, 0.8), "c": (0.5, 0.9, 1.0, 0.5, 0.5)}
LAMB_FAMILIES = ("struct_luma_s15", "struct_chroma_s30", "smooth_luma", "quant_chroma")
LAMB_BLOCKS = 40


# MSBD nets with trained-magnitude weights (synth.transplanted_msbd_state_dict): the (comp, qp) pairs whose transplant
# stays finite and O(10) at the output
TRANSPLANT_CASES = [("Luma", 32), ("Luma", 37), ("Chroma", 37)]
