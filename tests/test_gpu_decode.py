"""GPU parity of the integer kernels (post-process, map-to-partition decode, frame assembly, text formatting,
input prep) through the C ABI, bit-exact against the golden vectors frozen from the reference and against the
C oracle on fresh seeded inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import c_decode, decode_ref, nets_ref
from pmp_vvc_tip2023_b200 import Map2Partition, Metrics, ops, synth
from tests import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _unpack(g, name, n):
    hor = np.unpackbits(g[name + "_hor"], axis=1).reshape(n, 16, 16)
    ver = np.unpackbits(g[name + "_ver"], axis=1).reshape(n, 16, 16)
    return hor, ver, g[name + "_dire"]


def _decode(qt, bt, dire, cf):
    hor, ver, dout, flags = ops.map2partition(_cuda(qt.astype(np.uint8).reshape(-1, 64)), _cuda(bt), _cuda(dire), cf)
    return hor.cpu().numpy(), ver.cpu().numpy(), dout.cpu().numpy(), flags.cpu().numpy()


@pytest.mark.parametrize("name", sorted(cases.decode_cases().keys()))
def test_decode_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN, "decode_golden.npz"))
    qt, bt, dire, cf = cases.decode_cases()[name]
    hor, ver, dout = _unpack(g, name, qt.shape[0])
    h, v, d, flags = _decode(qt, bt, dire, cf)
    bad = [b for b in range(qt.shape[0])
           if not (np.array_equal(h[b], hor[b]) and np.array_equal(v[b], ver[b]) and np.array_equal(d[b], dout[b]))]
    # bit-exact except blocks the kernel itself flags as float32 near-ties of the reference's argmin
    unflagged = [b for b in bad if not (flags[b] & 1)]
    assert not unflagged, "%s: %d mismatching blocks not flagged near-tie: %s" % (name, len(unflagged), unflagged[:8])
    assert len(bad) <= max(1, qt.shape[0] // 50), "%s: %d near-tie mismatches" % (name, len(bad))


@pytest.mark.parametrize("chroma", [False, True])
@pytest.mark.parametrize("sigma", [0.0, 0.2, 0.45])
def test_decode_matches_c_oracle_large(chroma, sigma):
    n = 3000
    qt, bt, dire = synth.structured_maps(n, seed=900 + int(sigma * 100) + int(chroma), sigma=sigma, chroma=chroma)
    cf = 2 if chroma else 1
    hor, ver, dout = c_decode.map_to_partition_batch(qt, bt, dire, cf)
    h, v, d, flags = _decode(qt, bt, dire, cf)
    same = (h == hor).all(axis=(1, 2)) & (v == ver).all(axis=(1, 2)) & (d == dout).all(axis=(1, 2, 3))
    bad = np.nonzero(~same)[0]
    assert all(flags[b] & 1 for b in bad), "unflagged mismatches: %s" % bad[:8]
    assert len(bad) <= n // 200
    regions = flags >> 8
    assert regions.min() >= 0 and regions.max() <= 64


@pytest.mark.parametrize("tag", sorted(cases.LAMB_SETS))
def test_decode_nondefault_lamb_matches_reference_golden(tag):
    """The constructor thresholds lamb1..lamb5 (Map2Partition.py:100) are kernel arguments: goldens from the reference's
    Map_to_Partition(qt, bt, dire, cf, *lamb).get_partition(), through ops and through the drop-in class."""
    g = np.load(os.path.join(GOLDEN, "decode_lamb_golden.npz"))
    lamb = cases.LAMB_SETS[tag]
    allc = cases.decode_cases()
    for fam in cases.LAMB_FAMILIES:
        qt, bt, dire, cf = allc[fam]
        n = cases.LAMB_BLOCKS
        hor, ver, dout = _unpack(g, "%s_%s" % (tag, fam), n)
        h, v, d, flags = ops.map2partition(_cuda(qt[:n].astype(np.uint8).reshape(-1, 64)), _cuda(bt[:n]), _cuda(dire[:n]), cf,
                                           lamb=lamb)
        h, v, d, flags = h.cpu().numpy(), v.cpu().numpy(), d.cpu().numpy(), flags.cpu().numpy()
        bad = [b for b in range(n)
               if not (np.array_equal(h[b], hor[b]) and np.array_equal(v[b], ver[b]) and np.array_equal(d[b], dout[b]))]
        assert not [b for b in bad if not (flags[b] & 1)], (tag, fam, bad)
        assert len(bad) <= 1, (tag, fam, bad)
        par, dd = Map2Partition.Map_to_Partition(qt[3], bt[3], dire[3], cf, *lamb).get_partition()
        if 3 not in bad:
            assert np.array_equal(par[0][:16, :16], hor[3]) and np.array_equal(par[1][:16, :16], ver[3]) and np.array_equal(dd, dout[3])
    # the defaults passed explicitly are the default path
    qt, bt, dire, cf = allc["struct_luma_s15"]
    a = ops.map2partition(_cuda(qt.astype(np.uint8).reshape(-1, 64)), _cuda(bt), _cuda(dire), cf)
    b = ops.map2partition(_cuda(qt.astype(np.uint8).reshape(-1, 64)), _cuda(bt), _cuda(dire), cf, lamb=ops.DEFAULT_LAMB)
    assert all(torch.equal(x, y) for x, y in zip(a[:3], b[:3]))


@pytest.mark.parametrize("tol", [1e-2, 1e-3])
def test_near_threshold_flags(tol):
    """flags bits 1..3: a map value within `tol` of a decision threshold (bt: k+0.5, dire: +-0.5, pooled raw qt:
    0.5/1.5/2.5) -- the north star's "reported count of CTUs whose values fall within tolerance of a decision threshold"."""
    rng = np.random.default_rng(17)
    n = 400
    qt_raw = (rng.standard_normal((n, 1, 8, 8)) * 1.2 + 1.3).astype(np.float32)
    bt = np.round(rng.uniform(0, 3, (n, 3, 16, 16))).astype(np.float32) + rng.uniform(-0.2, 0.2, (n, 3, 16, 16)).astype(np.float32)
    dire = np.round(rng.uniform(-1, 1, (n, 3, 16, 16))).astype(np.float32) + rng.uniform(-0.2, 0.2, (n, 3, 16, 16)).astype(np.float32)
    # plant near-threshold values in known blocks (distance tol/2 and 2*tol)
    bt[0, 1, 3, 3] = 1.5 + tol / 2; bt[1, 0, 0, 0] = 2.5 - 2 * tol
    dire[2, 2, 5, 5] = -0.5 + tol / 2; dire[3, 0, 1, 1] = 0.5 + 2 * tol
    qt_raw[4, 0, 0:2, 0:2] = 1.5 - tol / 2; qt_raw[5, 0, 2:4, 2:4] = 3.2
    _, qu8 = ops.qt_postprocess(_cuda(qt_raw), want_f32=False)
    _, _, _, flags = ops.map2partition(qu8, _cuda(bt), _cuda(dire), 1, qt_raw=_cuda(qt_raw), near_tol=tol)
    f = flags.cpu().numpy()
    q = qt_raw.reshape(n, 4, 2, 4, 2).max(axis=(2, 4)).reshape(n, -1)
    want_q = (((q > 0) & (q < 3)) & (np.abs(q - np.floor(q) - 0.5) < np.float32(tol))).any(1)
    b = bt.reshape(n, -1)
    want_b = (np.abs(b - np.floor(b) - np.float32(0.5)) < np.float32(tol)).any(1)
    want_d = (np.abs(np.abs(dire.reshape(n, -1)) - np.float32(0.5)) < np.float32(tol)).any(1)
    assert np.array_equal((f & 2) != 0, want_b) and np.array_equal((f & 4) != 0, want_d) and np.array_equal((f & 8) != 0, want_q)
    assert f[0] & 2 and f[2] & 4 and f[4] & 8 and not (f[5] & 8)
    c = ops.flag_counts(flags)
    assert c["blocks"] == n and c["near_threshold_blocks"] == int((want_q | want_b | want_d).sum())
    # bits 4-6: the same tests at near_tol / 100
    t2 = np.float32(0.01) * np.float32(tol)
    tight_b = (np.abs(b - np.floor(b) - np.float32(0.5)) < t2).any(1)
    tight_d = (np.abs(np.abs(dire.reshape(n, -1)) - np.float32(0.5)) < t2).any(1)
    assert np.array_equal((f & 16) != 0, tight_b) and np.array_equal((f & 32) != 0, tight_d)
    assert c["near_threshold_blocks_tight"] <= c["near_threshold_blocks"]
    # without the raw qt map bit 3 stays clear; the plain entry point uses the handle's tolerance (1e-2)
    _, _, _, f2 = ops.map2partition(qu8, _cuda(bt), _cuda(dire), 1)
    assert not (f2.cpu().numpy() & 8).any()


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000])
def test_postprocess_ragged_sizes(n):
    """The coalesced kernel stages 128 blocks per CTA: ragged tails, both outputs, either output alone."""
    rng = np.random.default_rng(n)
    q = (rng.standard_normal((n, 1, 8, 8)) * 1.3 + 1.2).astype(np.float32)
    want = c_decode.qt_postprocess(q)
    of, ou = ops.qt_postprocess(_cuda(q))
    assert np.array_equal(of.cpu().numpy(), want) and np.array_equal(ou.cpu().numpy().reshape(-1, 1, 8, 8), want.astype(np.uint8))
    of2, none = ops.qt_postprocess(_cuda(q), want_u8=False)
    assert none is None and torch.equal(of2, of)


def test_decode_edge_cases():
    # empty batch, single block, ragged tail
    for n in (0, 1, 5):
        qt = np.zeros((n, 8, 8), np.float32)
        bt = np.zeros((n, 3, 16, 16), np.float32)
        dire = np.zeros((n, 3, 16, 16), np.float32)
        h, v, d, _ = _decode(qt, bt, dire, 1)
        assert h.shape == (n, 16, 16)
        if n:
            assert h[:, 0].all() and v[:, :, 0].all() and h[:, 1:].sum() == 0 and (d == 0).all()
    # NaN/inf maps must not hang or crash
    qt = np.zeros((4, 8, 8), np.float32)
    bt = np.full((4, 3, 16, 16), np.nan, np.float32)
    bt[1] = np.inf
    dire = np.full((4, 3, 16, 16), -np.inf, np.float32)
    _decode(qt, bt, dire, 2)
    torch.cuda.synchronize()


def test_map_to_parititon_dropin():
    qt, bt, dire, cf = cases.decode_cases()["struct_luma_s15"]
    for b in (0, 7, 31):
        want = decode_ref.map_to_partition(qt[b], bt[b], dire[b], cf)
        got = Map2Partition.map_to_parititon(qt[b], bt[b], dire[b], cf)
        for a, w in zip(got, want):
            assert a.dtype == w.dtype and np.array_equal(a, w)
    par, d = Map2Partition.Map_to_Partition(qt[0], bt[0], dire[0], cf).get_partition()
    assert par.shape == (2, 17, 17) and d.shape == (3, 16, 16)


def test_postprocess_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "postproc_golden.npz"))
    q = cases.postproc_inputs()
    of, ou = ops.qt_postprocess(_cuda(q))
    assert np.array_equal(of.cpu().numpy().astype(np.uint8), g["out"])
    assert np.array_equal(ou.cpu().numpy().reshape(-1, 1, 8, 8), g["out"])
    out = Metrics.eli_structual_error(_cuda(q))
    assert out.is_cuda and out.dtype == torch.float32 and np.array_equal(out.cpu().numpy().astype(np.uint8), g["out"])
    m = torch.tensor([[0., 1, 2, 3], [1, 1, 2, 2], [3, 3, 0, 0], [2, 1, 1, 1]])
    from oracle import postproc_ref
    assert np.array_equal(Metrics.check_square_unity(m).cpu().numpy(), postproc_ref.square_unity(m.numpy()))


def test_postprocess_random_vs_oracle():
    rng = np.random.default_rng(5)
    q = (rng.standard_normal((20000, 1, 8, 8)) * 1.3 + 1.2).astype(np.float32)
    q[::7] = np.round(q[::7] * 2) / 2
    _, ou = ops.qt_postprocess(_cuda(q), want_f32=False)
    assert np.array_equal(ou.cpu().numpy().reshape(-1, 1, 8, 8), c_decode.qt_postprocess(q).astype(np.uint8))


def test_sequence_file_matches_reference_file(tmp_path):
    """Decode the reference's own float maps -> assemble -> text: byte-identical to the file it wrote."""
    g = np.load(os.path.join(GOLDEN, "pipeline_golden.npz"))
    for comp in ("Luma", "Chroma"):
        want = open(os.path.join(GOLDEN, "pipeline_%s_QP32_PartitionMat.txt" % comp), "rb").read()
        path = str(tmp_path / (comp + ".txt"))
        Metrics.seq_post_process(torch.from_numpy(g[comp + "_qt"]).cuda(), g[comp + "_bt"], g[comp + "_dire"], comp,
                                 cases.PIPE_F, cases.PIPE_W, cases.PIPE_H, path)
        assert open(path, "rb").read() == want


def test_assemble_and_text_large():
    rng = np.random.default_rng(11)
    frames, bh, bw = 3, 16, 30                     # 1080p geometry
    n = frames * bh * bw
    hor = rng.integers(0, 2, (n, 16, 16), dtype=np.uint8)
    ver = rng.integers(0, 2, (n, 16, 16), dtype=np.uint8)
    qt = rng.integers(0, 4, (n, 8, 8), dtype=np.uint8)
    dire = rng.integers(-1, 2, (n, 3, 16, 16)).astype(np.int8)
    vals = ops.assemble_frames(_cuda(hor), _cuda(ver), _cuda(qt.reshape(n, 64)), _cuda(dire), frames, bh, bw)
    H = hor.reshape(frames, bh, bw, 16, 16).transpose(0, 1, 3, 2, 4).reshape(frames, bh * 16, bw * 16)
    V = ver.reshape(frames, bh, bw, 16, 16).transpose(0, 1, 3, 2, 4).reshape(frames, bh * 16, bw * 16)
    Q = qt.reshape(frames, bh, bw, 8, 8).transpose(0, 1, 3, 2, 4).reshape(frames, bh * 8, bw * 8)
    D = dire.reshape(frames, bh, bw, 3, 16, 16).transpose(0, 3, 1, 4, 2, 5).reshape(frames, 3, bh * 16, bw * 16)
    want = np.concatenate([H.reshape(frames, -1), V.reshape(frames, -1), Q.reshape(frames, -1),
                           D.reshape(frames, -1)], axis=1).astype(np.int8)
    assert vals.shape == (frames, 645120) and np.array_equal(vals.cpu().numpy(), want)
    text = ops.format_text(vals).cpu().numpy().tobytes()
    assert text == decode_ref.partition_text(H, V, Q, D)
    assert ops.format_text(vals[:0]).numel() == 0


@pytest.mark.parametrize("bits", [8, 10])
def test_cut_blocks_matches_oracle(bits):
    y, u, v = synth.synth_yuv420(416, 240, 2, seed=8, bitdepth=bits)
    if bits == 10:      # exercise the half-to-even ties and the 255 clip of np.round(y/4)
        y[0, :4, :8] = [[2, 6, 10, 14, 1018, 1022, 1023, 1021]] * 4
    by, bu, bv = nets_ref.cut_blocks(y, u, v, bits == 10)
    tv = lambda a: torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda()
    lb, cb = ops.cut_blocks(tv(y), tv(u), tv(v))
    assert np.array_equal(lb[:, 0].cpu().numpy(), by)
    want_c = nets_ref.chroma_net_input(by, bu, bv).numpy().astype(np.uint8)
    assert np.array_equal(cb.cpu().numpy(), want_c)


def test_cut_blocks_pipeline_golden():
    g = np.load(os.path.join(GOLDEN, "pipeline_golden.npz"))
    y, u, v = cases.pipeline_frames()
    tv = lambda a: torch.from_numpy(a.view(np.int16)).cuda()
    lb, cb = ops.cut_blocks(tv(y), tv(u), tv(v))
    assert np.array_equal(lb[:, 0].cpu().numpy(), g["by"])
    assert np.array_equal(cb[:, 1].cpu().numpy(), g["bu"]) and np.array_equal(cb[:, 2].cpu().numpy(), g["bv"])


def test_abi_error_paths():
    """Error behaviour of the C ABI: negative status + message, no exceptions, no crash (GPU present)."""
    import ctypes
    from pmp_vvc_tip2023_b200 import _lib
    h = _lib.Handle.get(0)
    L = _lib.lib()
    buf = torch.zeros(64, device="cuda")
    # unknown weight set
    rc = L.pmp_forward_q(h.ptr, 987654, buf.data_ptr(), _lib.IN_F32, 1, buf.data_ptr(), None)
    assert rc == -4 and b"unknown weight set" in L.pmp_last_error()
    # wrong tensor count / sizes for a net
    arr = np.zeros(10, np.float32)
    ptrs = (ctypes.c_void_p * 1)(arr.ctypes.data)
    numel = (ctypes.c_int64 * 1)(10)
    wset = ctypes.c_int(-1)
    assert L.pmp_weights_create(h.ptr, 0, ptrs, numel, 1, ctypes.byref(wset)) == -1
    with pytest.raises(_lib.PmpError):
        h.weights_create("Luma_Q", [np.zeros(3, np.float32)] * 20)
    # null pointers, bad chroma factor, negative batch
    assert L.pmp_qt_postprocess(h.ptr, None, 4, None, None, None) == -1
    assert L.pmp_map2partition(h.ptr, buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 1, 3, buf.data_ptr(), buf.data_ptr(),
                               buf.data_ptr(), None, None) == -1
    assert L.pmp_forward_q(h.ptr, 1, buf.data_ptr(), _lib.IN_F32, -1, buf.data_ptr(), None) == -1
    # empty batches are no-ops
    assert L.pmp_forward_q(h.ptr, 987654, None, _lib.IN_F32, 0, None, None) == 0
    assert L.pmp_assemble_frames(h.ptr, None, None, None, None, 0, 3, 4, None, None) == 0
    # a Q weight set is refused by the MSBD entry point
    sd = synth.seeded_state_dict("Luma_Q", 1)
    wq = h.weights_create("Luma_Q", list(sd.values()))
    x = torch.zeros((1, 1, 68, 68), device="cuda")
    o = torch.zeros((1, 2, 16, 16), device="cuda")
    rc = L.pmp_forward_msbd(h.ptr, wq, x.data_ptr(), _lib.IN_F32, buf.data_ptr(), 1, o.data_ptr(), o.data_ptr(), o.data_ptr(), None)
    assert rc == -4
    h.weights_destroy(wq)
    with pytest.raises(_lib.PmpError):
        h.weights_destroy(wq)
