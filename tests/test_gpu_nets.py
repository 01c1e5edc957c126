"""GPU parity of the four Down-Up-CNN forwards through the drop-in modules / C ABI against the golden outputs of
the reference's own PyTorch modules (Q nets: the reference's trained weights; MSBD nets: seeded weights, the
trained *_BD_*.pkl being absent from the mount) and the end-to-end PartitionMat file."""
import os

import numpy as np
import pytest
import torch

from oracle import nets_ref
from pmp_vvc_tip2023_b200 import _lib, Inference_QBD, Metrics, Model_QBD, ops, synth
from pmp_vvc_tip2023_b200.pipeline import PartitionPredictor
from pmp_vvc_tip2023_b200.weights import load_reference_pkl
from tests import cases

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# north_star tolerance: max-abs <= 1e-2 on map values vs the reference CPU fp32 model.  The exact-fp32 SIMT
# engine is held to summation-order noise instead.
TOL = {"simt": 2e-3, "tc": 1e-2}


def _nets(comp, qp):
    netq = getattr(Model_QBD, comp + "_Q_Net")()
    Inference_QBD.load_pretrain_model(netq, os.path.join(ROOT, "trained_models", "%s_Q_%d.pkl" % (comp, qp)))
    netb = getattr(Model_QBD, comp + "_MSBD_Net")()
    sdb = synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, qp))
    netb.load_state_dict({k: torch.from_numpy(v) for k, v in sdb.items()})
    return netq.cuda(), netb.cuda()


def _inputs(g, comp):
    by, bu, bv = g["by"], g["bu"], g["bv"]
    if comp == "Luma":
        return torch.from_numpy(by.astype(np.float32)).unsqueeze(1)
    return nets_ref.chroma_net_input(by, bu, bv)


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("comp", ["Luma", "Chroma"])
@pytest.mark.parametrize("qp", [22, 27, 32, 37])
def test_nets_match_reference_golden(engine, comp, qp):
    g = np.load(os.path.join(GOLDEN, "nets_golden.npz"))
    _lib.Handle.get(0).set_engine(_lib.ENGINE_TC if engine == "tc" else _lib.ENGINE_SIMT)
    x = _inputs(g, comp).cuda()
    netq, netb = _nets(comp, qp)
    qt = netq(x)
    want_qt = torch.from_numpy(g["%s_%d_qt" % (comp, qp)])
    err_q = float((qt.cpu() - want_qt).abs().max())
    assert qt.shape == (8, 1, 8, 8) and err_q <= TOL[engine], "qt max-abs %g" % err_q
    # feed the reference's qt so the MSBD check is independent of the Q check
    outs = netb(x, want_qt.cuda())
    want = torch.from_numpy(g["%s_%d_bd" % (comp, qp)])           # [N,3,2,16,16]
    err_b = max(float((outs[k].cpu() - want[:, k]).abs().max()) for k in range(3))
    assert err_b <= TOL[engine], "msbd max-abs %g" % err_b
    # uint8 pixels give the same result as float pixels
    qt8 = netq(x.to(torch.uint8))
    assert torch.equal(qt8, qt)


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_inference_pre_qbd_and_dataparallel(engine):
    from torch.utils.data import DataLoader, TensorDataset
    g = np.load(os.path.join(GOLDEN, "pipeline_golden.npz"))
    _lib.Handle.get(0).set_engine(_lib.ENGINE_TC if engine == "tc" else _lib.ENGINE_SIMT)
    for comp in ("Luma", "Chroma"):
        x = _inputs(g, comp)
        netq, netb = _nets(comp, 32)
        loader = DataLoader(TensorDataset(x), batch_size=5, shuffle=False)           # ragged last batch
        qt, bt, dire = Metrics.inference_pre_QBD(loader, torch.nn.DataParallel(netq).cuda(),
                                                 torch.nn.DataParallel(netb).cuda())
        assert not qt.is_cuda and qt.shape == (12, 1, 8, 8) and bt.shape == (12, 3, 16, 16)
        for got, key in ((qt, "_qt"), (bt, "_bt"), (dire, "_dire")):
            err = float((got - torch.from_numpy(g[comp + key])).abs().max())
            assert err <= TOL[engine], (comp, key, err)


@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_pipeline_file_matches_reference(engine, tmp_path):
    """frames -> cut -> nets -> post-process -> decode -> text, against the file the reference wrote.  Integer output
    is bit-exact except blocks whose float maps sit within tolerance of a decision threshold (counted)."""
    y, u, v = cases.pipeline_frames()
    pp = PartitionPredictor(0, engine=engine, chunk=7)          # chunk smaller than the block count: ragged chunks
    for comp in ("Luma", "Chroma"):
        pp.load_state_dicts(comp, 32, load_reference_pkl(os.path.join(ROOT, "trained_models", "%s_Q_32.pkl" % comp)),
                            synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, 32)))
    res = pp.predict_frames(y, u, v, qps=(32,), want_maps=True)
    g = np.load(os.path.join(GOLDEN, "pipeline_golden.npz"))
    for comp in ("Luma", "Chroma"):
        vals, qt, bt, dire, flags = res[(comp, 32)]
        for got, key in ((qt, "_qt"), (bt, "_bt"), (dire, "_dire")):
            err = float((got.cpu() - torch.from_numpy(g[comp + key])).abs().max())
            assert err <= TOL[engine], (comp, key, err)
        want = np.loadtxt(os.path.join(GOLDEN, "pipeline_%s_QP32_PartitionMat.txt" % comp), dtype=np.int64)
        got = vals.cpu().numpy().reshape(-1).astype(np.int64)
        assert got.shape == want.shape
        # per block: a differing block is excused only if ITS OWN golden maps hold a value closer to a decision threshold
        # than twice this block's measured map error, or its own argmin was flagged as a float32 near-tie
        nblk = qt.shape[0]
        bad_blocks = np.unique(cases.value_block_ids(cases.PIPE_F, cases.PIPE_H // 64, cases.PIPE_W // 64)[got != want])
        if bad_blocks.size:
            err_blk = np.max([np.abs(a.cpu().numpy().reshape(nblk, -1) - g[comp + key].reshape(nblk, -1)).max(axis=1)
                              for a, key in ((qt, "_qt"), (bt, "_bt"), (dire, "_dire"))], axis=0)
            dist = cases.threshold_distance(g[comp + "_qt"], g[comp + "_bt"], g[comp + "_dire"])
            tie = (flags.cpu().numpy() & 1) != 0
            unexcused = [int(b) for b in bad_blocks if not (dist[b] <= 2 * err_blk[b] + 1e-7 or tie[b])]
            assert not unexcused, "%s: blocks %s differ from the reference file without a near-threshold value" % (comp, unexcused)
            assert bad_blocks.size <= max(1, nblk // 50), "%s: %d of %d blocks differ" % (comp, bad_blocks.size, nblk)
        path = str(tmp_path / "f.txt")
        pp.write_partition_file(vals, path)
        assert np.array_equal(np.loadtxt(path, dtype=np.int64), got)


@pytest.mark.parametrize("engine", ["simt", "tc"])
@pytest.mark.parametrize("comp,qp", cases.TRANSPLANT_CASES)
def test_msbd_with_trained_magnitude_weights(engine, comp, qp):
    """The trained *_BD_*.pkl are absent offline; this pins the MSBD kernels on weights of TRAINED magnitude instead
    (trained Q-net tensors transplanted into the MSBD convs of the same shape: |activations| up to ~7e3, outputs of O(10)
    to O(50)) against the unmodified reference's outputs, at the north-star bar, and checks the fp16 range guard."""
    g = np.load(os.path.join(GOLDEN, "nets_transplant_golden.npz"))
    gb = np.load(os.path.join(GOLDEN, "nets_golden.npz"))
    h = _lib.Handle.get(0)
    h.set_engine(_lib.ENGINE_TC if engine == "tc" else _lib.ENGINE_SIMT)
    sat0 = h.saturation_count()
    x = _inputs(gb, comp).cuda()
    netb = getattr(Model_QBD, comp + "_MSBD_Net")()
    sdb = synth.transplanted_msbd_state_dict(comp, qp, os.path.join(ROOT, "trained_models"))
    netb.load_state_dict({k: torch.from_numpy(v) for k, v in sdb.items()})
    outs = netb.cuda()(x, torch.from_numpy(g["%s_%d_qt" % (comp, qp)]).cuda())
    want = torch.from_numpy(g["%s_%d_bd" % (comp, qp)])
    err = max(float((outs[k].cpu() - want[:, k]).abs().max()) for k in range(3))
    scale = float(want.abs().max())
    print("%s %s QP%d: max-abs %.2e on outputs up to %.1f" % (engine, comp, qp, err, scale))
    # tc: the absolute north-star bar although the outputs are 5-15x larger than real depth maps; simt: fp32 summation noise
    assert err <= (1e-2 if engine == "tc" else 2e-3 * max(1.0, scale / 4.0)), (err, scale)
    assert h.saturation_count() == sat0


def test_engines_agree_on_large_batch():
    """Size-independent property at a batch the oracle cannot reach: TC and SIMT engines agree within tolerance, and
    per-block results do not depend on batch composition."""
    by, bu, bv = synth.synth_blocks(600, seed=3)
    x = torch.from_numpy(by).unsqueeze(1).cuda()
    netq, netb = _nets("Luma", 27)
    h = _lib.Handle.get(0)
    h.set_engine(_lib.ENGINE_SIMT)
    q0 = netq(x)
    o0 = netb(x, q0)
    h.set_engine(_lib.ENGINE_TC)
    q1 = netq(x)
    o1 = netb(x, q0)
    assert float((q0 - q1).abs().max()) <= 1e-2
    assert max(float((a - b).abs().max()) for a, b in zip(o0, o1)) <= 1e-2
    q2 = netq(x[137:138])
    assert torch.equal(q2, q1[137:138])
    # odd batch: the CTA-pair conv kernel's peer CTA recomputes the last image and must drop it
    q3 = netq(x[:5])
    assert torch.equal(q3, q1[:5])
    o3 = netb(x[:5], q0[:5])
    assert all(torch.equal(a, b[:5]) for a, b in zip(o3, o1))


@pytest.mark.parametrize("cls", ["D", "C", "B", "A"])
def test_mixed_class_shapes_integer_path(cls):
    """BASELINE configs[4]: one frame of each VVC class shape through the whole pipeline.  The integer outputs must equal
    the C-oracle decode of the GPU's own float maps bit for bit (blocks flagged as float32 near-ties excepted)."""
    from oracle import c_decode
    w, h = synth.CLASS_SHAPES[cls]
    y, u, v = synth.synth_yuv420(w, h, 1, seed=40 + ord(cls))
    pp = PartitionPredictor(0, engine="tc", chunk=700)
    for comp in ("Luma", "Chroma"):
        pp.load_state_dicts(comp, 37, load_reference_pkl(os.path.join(ROOT, "trained_models", "%s_Q_37.pkl" % comp)),
                            synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, 37)))
    res = pp.predict_frames(y, u, v, qps=(37,), want_maps=True)
    bh, bw = h // 64, w // 64
    for comp in ("Luma", "Chroma"):
        vals, qt, bt, dire, flags = res[(comp, 37)]
        assert vals.shape == (1, ops.frame_values(bh, bw)) and qt.shape[0] == bh * bw
        assert torch.isfinite(qt).all() and torch.isfinite(bt).all() and torch.isfinite(dire).all()
        qi = c_decode.qt_postprocess(qt.cpu().numpy())
        hor, ver, dout = c_decode.map_to_partition_batch(qi[:, 0], bt.cpu().numpy(), dire.cpu().numpy(),
                                                         1 if comp == "Luma" else 2)
        R, C = bh * 16, bw * 16
        got = vals.cpu().numpy().reshape(-1)
        H = hor.reshape(bh, bw, 16, 16).transpose(0, 2, 1, 3).reshape(-1)
        V = ver.reshape(bh, bw, 16, 16).transpose(0, 2, 1, 3).reshape(-1)
        Q = qi[:, 0].astype(np.int8).reshape(bh, bw, 8, 8).transpose(0, 2, 1, 3).reshape(-1)
        D = dout.reshape(bh, bw, 3, 16, 16).transpose(2, 0, 3, 1, 4).reshape(-1)
        want = np.concatenate([H, V, Q, D]).astype(np.int8)
        diff = np.nonzero(got != want)[0]
        if diff.size:
            fl = flags.cpu().numpy()
            assert (fl & 1).any(), "%s %s: %d differing values and no near-tie flag" % (cls, comp, diff.size)
            assert diff.size <= 3 * 768 * int((fl & 1).sum())
        assert set(np.unique(got)) <= {-1, 0, 1, 2, 3}
        assert got[:R * C].reshape(R, C)[::16].all()          # every block's top row is an edge


def test_cli_driver_writes_reference_files(tmp_path):
    """The mirrored Inference_QBD driver (same flags / output tree as the reference, Inference_QBD.py:152-255) on a tiny
    10-bit sequence: the PartitionMat files must equal the ones the reference wrote for the same frames and weights."""
    import argparse
    y, u, v = cases.pipeline_frames()
    inp = tmp_path / "in"
    inp.mkdir()
    with open(inp / "pipe_192x128_10bit.yuv", "wb") as fp:
        for f in range(cases.PIPE_F):
            fp.write(y[f].tobytes()); fp.write(u[f].tobytes()); fp.write(v[f].tobytes())
    (tmp_path / "seqs.txt").write_text("pipe,pipe_192x128_10bit.yuv,192,128,%d,30\n#end!!!!\n" % cases.PIPE_F)
    cfg = tmp_path / "cfg"
    cfg.mkdir()
    (cfg / "pipe.cfg").write_text("InputFile : %s # comment\nInputBitDepth : 10\n" % (inp / "pipe_192x128_10bit.yuv"))
    args = Inference_QBD.build_parser().parse_args([
        "--jobID", "j1", "--inputDir", str(inp), "--outDir", str(tmp_path / "out"), "--batchSize", "5", "--startSeqID", "0",
        "--seqNum", "1", "--seqInfo", str(tmp_path / "seqs.txt"), "--cfgDir", str(cfg),
        "--modelDir", os.path.join(ROOT, "trained_models"), "--ssRatio", "1", "--missingBD", "seeded",
        "--gpus", str(min(2, torch.cuda.device_count())),           # 2 GPUs: one frame each, segments concatenated
        "--binaryOut"])
    Inference_QBD.inference_VVC_seqs(args)
    out = tmp_path / "out" / "j1" / "PartitionMat"
    names = sorted(n for n in os.listdir(out) if n.endswith(".txt"))
    assert len(names) == 8 and "pipe_192x128_10bit_Luma_QP32_PartitionMat.txt" in names
    from pmp_vvc_tip2023_b200 import partition_io
    for comp in ("Luma", "Chroma"):
        got = open(out / ("pipe_192x128_10bit_%s_QP32_PartitionMat.txt" % comp), "rb").read()
        want = open(os.path.join(GOLDEN, "pipeline_%s_QP32_PartitionMat.txt" % comp), "rb").read()
        assert got == want
        # --binaryOut: the raw int8 form for the VTM-side reader holds the same values in the same order
        vals, rows, cols = partition_io.read_partition_bin(str(out / ("pipe_192x128_10bit_%s_QP32_PartitionMat.bin" % comp)))
        assert (rows, cols) == (16 * (cases.PIPE_H >> 6), 16 * (cases.PIPE_W >> 6))
        assert np.array_equal(vals, partition_io.text_to_values(os.path.join(GOLDEN, "pipeline_%s_QP32_PartitionMat.txt" % comp),
                                                                cases.PIPE_H, cases.PIPE_W))
    assert os.path.exists(tmp_path / "out" / "j1" / "Time_Sta_0_1.txt")


@pytest.mark.parametrize("comp,qp", [("Luma", 37), ("Chroma", 22)])
def test_bf16_split_engine_within_tolerance(comp, qp):
    """PMP_TC_BF16 operands (hi+lo bf16, ~16 mantissa bits): still inside the 1e-2 bar (bf16x3 ~1e-3)."""
    g = np.load(os.path.join(GOLDEN, "nets_golden.npz"))
    h = _lib.Handle.get(0)
    h.set_engine(_lib.ENGINE_TC, _lib.TC_BF16)
    try:
        x = _inputs(g, comp).cuda()
        netq, netb = _nets(comp, qp)
        want_qt = torch.from_numpy(g["%s_%d_qt" % (comp, qp)])
        err_q = float((netq(x).cpu() - want_qt).abs().max())
        outs = netb(x, want_qt.cuda())
        want = torch.from_numpy(g["%s_%d_bd" % (comp, qp)])
        err_b = max(float((outs[k].cpu() - want[:, k]).abs().max()) for k in range(3))
        assert err_q <= 1e-2 and err_b <= 1e-2, (err_q, err_b)
    finally:
        h.set_engine(_lib.ENGINE_TC, _lib.TC_FP16)


def test_predict_frames_host_out_matches_device_results():
    """The public API's overlapped device->host copies (side stream, one per finished component) deliver exactly the
    vectors the call returns on the device, for several QPs and repeated calls into the same pinned buffers."""
    y, u, v = cases.pipeline_frames()
    pp = PartitionPredictor(0, engine="tc", chunk=5)
    qps = (27, 32)
    for comp in ("Luma", "Chroma"):
        for qp in qps:
            pp.load_state_dicts(comp, qp, load_reference_pkl(os.path.join(ROOT, "trained_models", "%s_Q_%d.pkl" % (comp, qp))),
                                synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, qp)))
    ref = pp.predict_frames(y, u, v, qps=qps)
    host = {k: torch.empty(t.shape, dtype=torch.int8).pin_memory() for k, t in ref.items()}
    for _ in range(2):
        for t in host.values():
            t.fill_(-7)
        res = pp.predict_frames(y, u, v, qps=qps, host_out=host)
        pp.synchronize()
        for k in ref:
            assert torch.equal(res[k].cpu(), ref[k].cpu())
            assert torch.equal(host[k], ref[k].cpu()), k


# ---------------------------------------------------------------------------------------------------------------------
# Parity at the bench configuration: the oracle (oracle.nets_ref = the reference's forwards restated in plain torch, CPU
# fp32, pinned to the reference by tests/golden) run LIVE on this host against the TC engine on one full 1080p frame
# (480 blocks) per net, and on 256 textured blocks at the worst-case QPs of SURVEY 7.3 -- the precision tail sits in ~4 %
# knife-edge blocks that 8 golden blocks cannot see.  Bar: max-abs <= 1e-2 (north star); the measured value is printed.
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_and_gpu(comp, qp, by, bu, bv, engine="tc"):
    luma = comp == "Luma"
    x = torch.from_numpy(by.astype(np.float32)).unsqueeze(1) if luma else nets_ref.chroma_net_input(by, bu, bv)
    sdq = load_reference_pkl(os.path.join(ROOT, "trained_models", "%s_Q_%d.pkl" % (comp, qp)))
    sdb_np = synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, qp))
    sdb = {k: torch.from_numpy(v) for k, v in sdb_np.items()}
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    want_qt, want_bt, want_dire = nets_ref.predict_maps(sdq, sdb, x, luma, batch=120)
    pp = PartitionPredictor(0, engine=engine, chunk=200)            # ragged chunks: 200 + 200 + 80
    pp.load_state_dicts(comp, qp, sdq, sdb_np)
    wq, wb = pp._wsets[(comp, qp)]
    xu8 = x.to(torch.uint8).cuda()
    qt, bt, dire = ops.predict_maps(wq, wb, xu8, handle=pp.handle)
    torch.cuda.synchronize()
    errs = {"qt": float((qt.cpu() - want_qt).abs().max()), "bt": float((bt.cpu() - want_bt).abs().max()),
            "dire": float((dire.cpu() - want_dire).abs().max())}
    sat = pp.handle.saturation_count()
    pp.close()
    return errs, sat, (want_qt, want_bt, want_dire), (qt.cpu(), bt.cpu(), dire.cpu())


@pytest.mark.parametrize("comp,qp", [("Luma", 22), ("Luma", 37), ("Chroma", 22), ("Chroma", 32)])
def test_full_1080p_frame_matches_live_oracle(comp, qp):
    y, u, v = synth.synth_yuv420(1920, 1080, 1, seed=300 + qp)
    by, bu, bv = nets_ref.cut_blocks(y, u, v, True)
    assert by.shape[0] == 480
    errs, sat, want, got = _oracle_and_gpu(comp, qp, by, bu, bv)
    print("\n[parity] %s QP%d, 480 blocks (one 1080p frame), TC engine vs live CPU-fp32 oracle: max-abs qt %.3e bt %.3e dire %.3e"
          % (comp, qp, errs["qt"], errs["bt"], errs["dire"]))
    assert sat == 0, "fp16 saturation events: %d" % sat
    assert max(errs.values()) <= 1e-2, errs
    # integer path on top: post-processed qt equal except blocks within the measured error of a rounding threshold
    qi_want = c_decode_qt(want[0].numpy())
    qi_got = c_decode_qt(got[0].numpy())
    diff = np.nonzero((qi_want != qi_got).reshape(480, -1).any(1))[0]
    dist = cases.threshold_distance(want[0].numpy(), np.zeros((480, 1), np.float32), np.zeros((480, 1), np.float32))   # qt only
    assert all(dist[b] <= 2 * errs["qt"] + 1e-7 for b in diff), diff


def c_decode_qt(q):
    from oracle import c_decode
    return c_decode.qt_postprocess(q)


@pytest.mark.parametrize("qp", [22, 37])
def test_textured_blocks_worst_case_qps(qp):
    """>= 256 textured blocks on the Luma Q + MSBD nets at QP 22 / 37 (the survey's worst cases for reduced precision)."""
    by, bu, bv = synth.synth_blocks(256, seed=50 + qp)
    by8 = np.clip(by, 0, 255).astype(np.uint8)
    errs, sat, _, _ = _oracle_and_gpu("Luma", qp, by8, None, None)
    print("\n[parity] Luma QP%d, 256 textured blocks: max-abs qt %.3e bt %.3e dire %.3e" % (qp, errs["qt"], errs["bt"], errs["dire"]))
    assert sat == 0 and max(errs.values()) <= 1e-2, errs


def test_predictors_do_not_share_engine_state():
    """Two predictors on one device keep their own engine (each owns a handle): a simt-vs-tc comparison really compares
    two engines, and neither disturbs the Model_QBD modules' default handle."""
    y, u, v = cases.pipeline_frames()
    h0 = _lib.Handle.get(0)
    h0.set_engine(_lib.ENGINE_TC)
    a = PartitionPredictor(0, engine="tc", chunk=6)
    b = PartitionPredictor(0, engine="simt", chunk=6)
    assert a.handle.engine() == _lib.ENGINE_TC and b.handle.engine() == _lib.ENGINE_SIMT and h0.engine() == _lib.ENGINE_TC
    for pp in (a, b):
        pp.load_state_dicts("Luma", 32, load_reference_pkl(os.path.join(ROOT, "trained_models", "Luma_Q_32.pkl")),
                            synth.seeded_state_dict("Luma_MSBD", cases.msbd_seed("Luma", 32)))
    ra = a.predict_frames(y, u, v, qps=(32,), comps=("Luma",), want_maps=True)[("Luma", 32)]
    l0 = b.handle.launch_count()
    rb = b.predict_frames(y, u, v, qps=(32,), comps=("Luma",), want_maps=True)[("Luma", 32)]
    assert b.handle.launch_count() > l0 and a.handle.engine() == _lib.ENGINE_TC
    d = float((ra[2] - rb[2]).abs().max())
    assert 0.0 < d <= 1e-2, d            # different engines: close, but not the same bits
    c = a.counts()
    assert c["total"]["blocks"] == 12 and set(c["per_component_qp"]) == {"Luma_QP32"}
    a.close(); b.close()


def test_fp16_saturation_is_reported():
    """Activations beyond +-65504 cannot be split into fp16 hi/lo: the engine counts them instead of clamping silently."""
    sdq = {k: torch.from_numpy(v) for k, v in synth.seeded_state_dict("Luma_Q", 3).items()}
    big = {k: (v * 2000.0 if k.startswith("resblock_q1.left.0") else v) for k, v in sdq.items()}
    x = torch.from_numpy(np.clip(synth.synth_blocks(4, seed=1)[0], 0, 255).astype(np.float32)).unsqueeze(1).cuda()
    h = _lib.Handle.get(0)
    h.set_engine(_lib.ENGINE_TC)
    h.saturation_count(reset=True)
    for sd, expect in ((sdq, False), (big, True)):
        net = Model_QBD.Luma_Q_Net()
        net.load_state_dict(sd)
        net.cuda()(x)
        n = h.saturation_count(reset=True)
        assert (n > 0) == expect, n
