"""Freeze golden vectors from the UNMODIFIED reference (build container only).

Run:  python oracle/gen_golden.py          (needs /root/reference; writes tests/golden/)

The reference is Python, so it cannot travel to the GPU box as a binary; instead
its outputs on deterministic inputs are frozen here and the oracle restatements
(oracle/*.py, oracle/decode_ref.c) plus the CUDA path are checked against them.
Shims: torch.Tensor.cuda -> identity (Metrics.py hard-codes
.cuda()), torch.load(map_location='cpu') for the CUDA-tagged legacy pickles.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
torch.Tensor.cuda = lambda self, *a, **k: self        # shim: no GPU in the build container

import Map2Partition as RefM2P      # noqa: E402
import Metrics as RefMetrics        # noqa: E402
import Model_QBD as RefModel        # noqa: E402

from pmp_vvc_tip2023_b200 import synth  # noqa: E402
from tests import cases             # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_ref_sd(path):
    sd = torch.load(path, map_location="cpu", weights_only=False)
    return {k.split("module.", 1)[-1] if k.startswith("module.") else k: v for k, v in sd.items()}


def gen_decode():
    res = {}
    for name, (qt, bt, dire, cf) in cases.decode_cases().items():
        n = qt.shape[0]
        hor = np.zeros((n, 16, 16), np.uint8)
        ver = np.zeros((n, 16, 16), np.uint8)
        dout = np.zeros((n, 3, 16, 16), np.int8)
        for b in range(n):
            h, v, d = RefM2P.map_to_parititon(qt[b], bt[b], dire[b], cf)
            hor[b], ver[b], dout[b] = h, v, d
        res[name + "_hor"] = np.packbits(hor.reshape(n, -1), axis=1)
        res[name + "_ver"] = np.packbits(ver.reshape(n, -1), axis=1)
        res[name + "_dire"] = dout
        res[name + "_insum"] = np.array([float(np.abs(bt).sum()), float(np.abs(dire).sum()), float(qt.sum())])
        print("decode", name, n, "edges", hor.mean())
    np.savez_compressed(os.path.join(OUT, "decode_golden.npz"), **res)


def gen_decode_lamb():
    """Map_to_Partition with non-default constructor thresholds lamb1..lamb5 (Map2Partition.py:100)."""
    res = {}
    allc = cases.decode_cases()
    for tag, lamb in cases.LAMB_SETS.items():
        for fam in cases.LAMB_FAMILIES:
            qt, bt, dire, cf = allc[fam]
            n = cases.LAMB_BLOCKS
            hor = np.zeros((n, 16, 16), np.uint8)
            ver = np.zeros((n, 16, 16), np.uint8)
            dout = np.zeros((n, 3, 16, 16), np.int8)
            for b in range(n):
                par, d = RefM2P.Map_to_Partition(qt[b], bt[b], dire[b], cf, *lamb).get_partition()
                hor[b], ver[b], dout[b] = par[0][:16, :16], par[1][:16, :16], d
            res["%s_%s_hor" % (tag, fam)] = np.packbits(hor.reshape(n, -1), axis=1)
            res["%s_%s_ver" % (tag, fam)] = np.packbits(ver.reshape(n, -1), axis=1)
            res["%s_%s_dire" % (tag, fam)] = dout
            print("decode lamb", tag, fam, "edges", hor.mean())
    np.savez_compressed(os.path.join(OUT, "decode_lamb_golden.npz"), **res)


def gen_postproc():
    q = cases.postproc_inputs()
    out = RefMetrics.eli_structual_error(torch.from_numpy(q.copy())).numpy()
    np.savez_compressed(os.path.join(OUT, "postproc_golden.npz"), out=out.astype(np.uint8),
                        insum=np.array([float(np.abs(q).sum())]))
    print("postproc", q.shape)


@torch.no_grad()
def gen_nets():
    by, bu, bv = cases.net_blocks()
    xl = torch.from_numpy(by.astype(np.float32)).unsqueeze(1)
    xc = torch.cat([torch.nn.functional.max_pool2d(xl, 2),
                    torch.from_numpy(bu.astype(np.float32)).unsqueeze(1),
                    torch.from_numpy(bv.astype(np.float32)).unsqueeze(1)], 1)
    res = {"by": by, "bu": bu, "bv": bv}
    for comp, x in (("Luma", xl), ("Chroma", xc)):
        for qp in (22, 27, 32, 37):
            netq = getattr(RefModel, comp + "_Q_Net")()
            netq.load_state_dict(load_ref_sd(os.path.join(REF, "trained_models", "%s_Q_%d.pkl" % (comp, qp))))
            netb = getattr(RefModel, comp + "_MSBD_Net")()
            sdb = synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, qp))
            netb.load_state_dict({k: torch.from_numpy(v) for k, v in sdb.items()})
            netq.eval(), netb.eval()
            qt = netq(x)
            o0, o1, o2 = netb(x, qt)
            res["%s_%d_qt" % (comp, qp)] = qt.numpy()
            res["%s_%d_bd" % (comp, qp)] = torch.stack([o0, o1, o2], 1).numpy()     # [N,3,2,16,16]
            print("nets", comp, qp, float(qt.min()), float(qt.max()), float(o2.min()), float(o2.max()))
    np.savez_compressed(os.path.join(OUT, "nets_golden.npz"), **res)


@torch.no_grad()
def gen_nets_transplant():
    """MSBD nets of the UNMODIFIED reference with trained-magnitude weights (synth.transplanted_msbd_state_dict)."""
    by, bu, bv = cases.net_blocks()
    xl = torch.from_numpy(by.astype(np.float32)).unsqueeze(1)
    xc = torch.cat([torch.nn.functional.max_pool2d(xl, 2),
                    torch.from_numpy(bu.astype(np.float32)).unsqueeze(1),
                    torch.from_numpy(bv.astype(np.float32)).unsqueeze(1)], 1)
    res = {}
    for comp, qp in cases.TRANSPLANT_CASES:
        x = xl if comp == "Luma" else xc
        netq = getattr(RefModel, comp + "_Q_Net")()
        netq.load_state_dict(load_ref_sd(os.path.join(REF, "trained_models", "%s_Q_%d.pkl" % (comp, qp))))
        netb = getattr(RefModel, comp + "_MSBD_Net")()
        sdb = synth.transplanted_msbd_state_dict(comp, qp, os.path.join(REF, "trained_models"))
        netb.load_state_dict({k: torch.from_numpy(v) for k, v in sdb.items()})
        netq.eval(), netb.eval()
        qt = netq(x)
        o0, o1, o2 = netb(x, qt)
        res["%s_%d_qt" % (comp, qp)] = qt.numpy()
        res["%s_%d_bd" % (comp, qp)] = torch.stack([o0, o1, o2], 1).numpy()
        print("nets transplant", comp, qp, float(o2.min()), float(o2.max()))
    np.savez_compressed(os.path.join(OUT, "nets_transplant_golden.npz"), **res)


@torch.no_grad()
def gen_pipeline():
    """inference_pre_QBD + seq_post_process (file on disk) on a tiny 2-frame sequence."""
    from torch.utils.data import DataLoader, TensorDataset
    import tempfile
    import Inference_QBD as RefInf
    w, h, nf = cases.PIPE_W, cases.PIPE_H, cases.PIPE_F
    y, u, v = cases.pipeline_frames()
    with tempfile.NamedTemporaryFile(suffix=".yuv") as tf:
        for f in range(nf):
            tf.write(y[f].tobytes()); tf.write(u[f].tobytes()); tf.write(v[f].tobytes())
        tf.flush()
        by, bu, bv = RefInf.output_block_yuv(tf.name, w, h, block_size=64, in_overlap=4, numfrm=nf,
                                             SubSampleRatio=1, is10bit=True)
    xl = torch.from_numpy(by.astype(np.float32)).unsqueeze(1)
    xc = torch.cat([torch.nn.functional.max_pool2d(xl, 2),
                    torch.from_numpy(bu.astype(np.float32)).unsqueeze(1),
                    torch.from_numpy(bv.astype(np.float32)).unsqueeze(1)], 1)
    res = {"by": by, "bu": bu, "bv": bv}
    for comp, x in (("Luma", xl), ("Chroma", xc)):
        qp = 32
        netq = getattr(RefModel, comp + "_Q_Net")()
        netq.load_state_dict(load_ref_sd(os.path.join(REF, "trained_models", "%s_Q_%d.pkl" % (comp, qp))))
        netb = getattr(RefModel, comp + "_MSBD_Net")()
        sdb = synth.seeded_state_dict(comp + "_MSBD", cases.msbd_seed(comp, qp))
        netb.load_state_dict({k: torch.from_numpy(v) for k, v in sdb.items()})
        loader = DataLoader(TensorDataset(x), batch_size=5, shuffle=False)
        with contextlib.redirect_stdout(io.StringIO()):
            qt, bt, dire = RefMetrics.inference_pre_QBD(loader, netq, netb)
            path = os.path.join(OUT, "pipeline_%s_QP%d_PartitionMat.txt" % (comp, qp))
            RefMetrics.seq_post_process(qt.clone(), bt.numpy(), dire.numpy(), comp, nf, w, h, path)
        res[comp + "_qt"] = qt.numpy()
        res[comp + "_bt"] = bt.numpy()
        res[comp + "_dire"] = dire.numpy()
        print("pipeline", comp, os.path.getsize(path))
    np.savez_compressed(os.path.join(OUT, "pipeline_golden.npz"), **res)


def gen_demo_fixture():
    """First frame of two reference-produced demo PartitionMat files (format fixture)."""
    for comp in ("Luma", "Chroma"):
        src = os.path.join(REF, "codec", "demo", "PartitionMat", "RaceHorses_416x240_30_%s_QP32_PartitionMat.txt" % comp)
        lines = open(src, "rb").read().split(b"\n")
        per_frame = 24192                # 416x240: R=48, C=96 -> 2*4608 + 1152 + 13824
        with open(os.path.join(OUT, "demo_RaceHorses_%s_QP32_frame0.txt" % comp), "wb") as f:
            f.write(b"\n".join(lines[:per_frame]) + b"\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["decode", "decode_lamb", "postproc", "nets", "nets_transplant", "pipeline", "demo"]
    if "decode" in which:
        gen_decode()
    if "decode_lamb" in which:
        gen_decode_lamb()
    if "postproc" in which:
        gen_postproc()
    if "nets" in which:
        gen_nets()
    if "nets_transplant" in which:
        gen_nets_transplant()
    if "pipeline" in which:
        gen_pipeline()
    if "demo" in which:
        gen_demo_fixture()
