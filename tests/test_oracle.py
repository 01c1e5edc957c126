"""The oracle restatements (oracle/*.py, oracle/decode_ref.c) against golden vectors frozen
from the UNMODIFIED reference by oracle/gen_golden.py.  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import c_decode, decode_ref, nets_ref, postproc_ref
from pmp_vvc_tip2023_b200 import synth
from pmp_vvc_tip2023_b200.weights import load_reference_pkl
from tests import cases

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def decode_golden():
    return np.load(os.path.join(GOLDEN, "decode_golden.npz"))


def _unpack(g, name, n):
    hor = np.unpackbits(g[name + "_hor"], axis=1).reshape(n, 16, 16)
    ver = np.unpackbits(g[name + "_ver"], axis=1).reshape(n, 16, 16)
    return hor, ver, g[name + "_dire"]


@pytest.mark.parametrize("name", sorted(cases.decode_cases().keys()))
def test_decode_c_oracle_matches_reference(decode_golden, name):
    qt, bt, dire, cf = cases.decode_cases()[name]
    insum = decode_golden[name + "_insum"]
    assert np.allclose(insum, [np.abs(bt).sum(), np.abs(dire).sum(), qt.sum()], rtol=1e-6), \
        "case generator drifted from the frozen inputs"
    hor, ver, dout = _unpack(decode_golden, name, qt.shape[0])
    h, v, d = c_decode.map_to_partition_batch(qt, bt, dire, cf)
    assert np.array_equal(h, hor) and np.array_equal(v, ver) and np.array_equal(d, dout)


@pytest.mark.parametrize("name", ["struct_luma_s15", "quant_chroma", "wild_luma", "smooth_chroma"])
def test_decode_numpy_oracle_matches_reference(decode_golden, name):
    qt, bt, dire, cf = cases.decode_cases()[name]
    n = min(24, qt.shape[0])
    hor, ver, dout = _unpack(decode_golden, name, qt.shape[0])
    for b in range(n):
        h, v, d = decode_ref.map_to_partition(qt[b], bt[b], dire[b], cf)
        assert np.array_equal(h, hor[b]) and np.array_equal(v, ver[b]) and np.array_equal(d, dout[b]), (name, b)


@pytest.mark.parametrize("tag", sorted(cases.LAMB_SETS))
def test_decode_oracles_match_reference_nondefault_lamb(tag):
    """Map_to_Partition(..., lamb1..lamb5) with non-default constructor thresholds (Map2Partition.py:100)."""
    g = np.load(os.path.join(GOLDEN, "decode_lamb_golden.npz"))
    lamb = cases.LAMB_SETS[tag]
    allc = cases.decode_cases()
    for fam in cases.LAMB_FAMILIES:
        qt, bt, dire, cf = allc[fam]
        n = cases.LAMB_BLOCKS
        hor, ver, dout = _unpack(g, "%s_%s" % (tag, fam), n)
        h, v, d = c_decode.map_to_partition_batch(qt[:n], bt[:n], dire[:n], cf, lamb=lamb)
        assert np.array_equal(h, hor) and np.array_equal(v, ver) and np.array_equal(d, dout), (tag, fam)
        for b in range(0, n, 13):
            h1, v1, d1 = decode_ref.map_to_partition(qt[b], bt[b], dire[b], cf, lamb)
            assert np.array_equal(h1, hor[b]) and np.array_equal(v1, ver[b]) and np.array_equal(d1, dout[b]), (tag, fam, b)


def test_postproc_oracles_match_reference():
    g = np.load(os.path.join(GOLDEN, "postproc_golden.npz"))
    q = cases.postproc_inputs()
    assert np.isclose(np.abs(q).sum(), g["insum"][0], rtol=1e-6)
    assert np.array_equal(postproc_ref.eli_structural_error(q).astype(np.uint8), g["out"])
    assert np.array_equal(c_decode.qt_postprocess(q).astype(np.uint8), g["out"])


@pytest.mark.parametrize("comp", ["Luma", "Chroma"])
@pytest.mark.parametrize("qp", [22, 27, 32, 37])
def test_nets_oracle_matches_reference(comp, qp):
    g = np.load(os.path.join(GOLDEN, "nets_golden.npz"))
    by, bu, bv = cases.net_blocks()
    assert np.array_equal(by, g["by"]) and np.array_equal(bu, g["bu"]) and np.array_equal(bv, g["bv"])
    luma = comp == "Luma"
    x = torch.from_numpy(by.astype(np.float32)).unsqueeze(1) if luma else nets_ref.chroma_net_input(by, bu, bv)
    sdq = load_reference_pkl(os.path.join(ROOT, "trained_models", "%s_Q_%d.pkl" % (comp, qp)))
    sdb = {k: torch.from_numpy(v) for k, v in
           synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, qp)).items()}
    with torch.no_grad():
        qt = nets_ref.q_net_forward(sdq, x, luma)
        o = torch.stack(nets_ref.msbd_net_forward(sdb, x, qt, luma), 1)
    # same ops, same library: only thread-count dependent summation-order noise is allowed
    assert float((qt - torch.from_numpy(g["%s_%d_qt" % (comp, qp)])).abs().max()) < 2e-4
    ref = torch.from_numpy(g["%s_%d_bd" % (comp, qp)])
    assert float((o - ref).abs().max()) < 2e-3 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("comp,qp", cases.TRANSPLANT_CASES)
def test_msbd_oracle_matches_reference_with_trained_magnitude_weights(comp, qp):
    """MSBD oracle vs the unmodified reference with transplanted trained weights (activations of trained magnitude)."""
    g = np.load(os.path.join(GOLDEN, "nets_transplant_golden.npz"))
    by, bu, bv = cases.net_blocks()
    luma = comp == "Luma"
    x = torch.from_numpy(by.astype(np.float32)).unsqueeze(1) if luma else nets_ref.chroma_net_input(by, bu, bv)
    sdb = {k: torch.from_numpy(v) for k, v in
           synth.transplanted_msbd_state_dict(comp, qp, os.path.join(ROOT, "trained_models")).items()}
    with torch.no_grad():
        o = torch.stack(nets_ref.msbd_net_forward(sdb, x, torch.from_numpy(g["%s_%d_qt" % (comp, qp)]), luma), 1)
    ref = torch.from_numpy(g["%s_%d_bd" % (comp, qp)])
    assert float(ref.abs().max()) > 8.0                          # the case is what it claims: outputs of O(10)
    assert float((o - ref).abs().max()) < 2e-3 * float(ref.abs().max())


def test_pipeline_oracle_matches_reference_file():
    """cut_blocks -> nets -> postproc -> decode -> text == the file the reference wrote."""
    g = np.load(os.path.join(GOLDEN, "pipeline_golden.npz"))
    y, u, v = cases.pipeline_frames()
    by, bu, bv = nets_ref.cut_blocks(y, u, v, True)
    assert np.array_equal(by, g["by"]) and np.array_equal(bu, g["bu"]) and np.array_equal(bv, g["bv"])
    for comp in ("Luma", "Chroma"):
        luma = comp == "Luma"
        # decode the reference's own float maps: must reproduce its file byte for byte
        qt_i = postproc_ref.eli_structural_error(g[comp + "_qt"])[:, 0]
        hor, ver, qtm, dire = decode_ref.sequence_partition(
            qt_i, g[comp + "_bt"], g[comp + "_dire"], luma, cases.PIPE_F, cases.PIPE_W, cases.PIPE_H)
        want = open(os.path.join(GOLDEN, "pipeline_%s_QP32_PartitionMat.txt" % comp), "rb").read()
        assert decode_ref.partition_text(hor, ver, qtm, dire) == want


def test_demo_fixture_invariants():
    """Structural invariants of a reference-produced PartitionMat frame (format fixture)."""
    for comp in ("Luma", "Chroma"):
        vals = np.loadtxt(os.path.join(GOLDEN, "demo_RaceHorses_%s_QP32_frame0.txt" % comp), dtype=np.int64)
        R, C = (240 // 64) * 16, (416 // 64) * 16
        assert vals.size == 2 * R * C + (R // 2) * (C // 2) + 3 * R * C
        hor = vals[:R * C].reshape(R, C)
        ver = vals[R * C:2 * R * C].reshape(R, C)
        qt = vals[2 * R * C:2 * R * C + (R // 2) * (C // 2)].reshape(R // 2, C // 2)
        dire = vals[2 * R * C + (R // 2) * (C // 2):].reshape(3, R, C)
        assert set(np.unique(hor)) <= {0, 1} and set(np.unique(ver)) <= {0, 1}
        assert set(np.unique(qt)) <= {0, 1, 2, 3} and set(np.unique(dire)) <= {-1, 0, 1}
        assert hor[::16].all() and ver[:, ::16].all()       # every block's top row / left column is an edge
