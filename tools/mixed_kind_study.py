"""CPU accuracy study for the next step of the split-precision scheme: which operand formats the two
correction products a_hi*w_lo and a_lo*w_hi tolerate, and what Winograd F(2x2,3x3) on split operands (2.25x fewer MACs for
the 3x3 layers; "wino3-*": every 3x3 "same" conv through the transform, the rest direct fp16x3) costs in accuracy.  Emulates per-conv operand rounding with fp32 results on the trained Q
nets (the MSBD *.pkl are not in this checkout) over synthetic textured blocks and prints the max-abs error of the 8x8 QT map
against fp64 convolutions of the unrounded operands.  Pure PyTorch on the CPU; no library code involved.

    python tools/mixed_kind_study.py [n_blocks]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pmp_vvc_tip2023_b200 import synth  # noqa: E402
from pmp_vvc_tip2023_b200.weights import load_reference_pkl  # noqa: E402


def r16(x):
    return x.to(torch.float16).to(torch.float64)


def rbf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def r8(x, fmt, scaled=True):
    """Round to an 8-bit float (e4m3fn / e5m2); with `scaled` a per-tensor power-of-two scale puts max|x| just under the
    format's largest value (what a per-layer static scale would do)."""
    dt, top = (torch.float8_e4m3fn, 448.0) if fmt == "e4m3" else (torch.float8_e5m2, 57344.0)
    m = float(x.abs().max())
    s = 1.0
    if scaled and m > 0:
        s = 2.0 ** np.floor(np.log2(top / m))
    y = (x * s).clamp(-top, top).to(torch.float32).to(dt).to(torch.float64)
    return y / s


# Winograd F(2x2, 3x3) (Lavin & Gray): Y = A^T [(G g G^T) . (B^T d B)] A per 4x4 input tile, 16 multiplies per 4 outputs
_BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=torch.float64)
_G = torch.tensor([[1, 0, 0], [0.5, 0.5, 0.5], [0.5, -0.5, 0.5], [0, 0, 1]], dtype=torch.float64)
_AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=torch.float64)


def winograd3x3(x, w, rnd):
    """3x3 'same' conv through F(2x2,3x3) with both transformed operands split into hi + lo (rnd) and three products."""
    B, C, H, W = x.shape
    co = w.shape[0]
    tiles = F.unfold(F.pad(x, (1, 1, 1, 1)), kernel_size=4, stride=2).view(B, C, 16, -1)          # [B, C, 16, T]
    V = torch.einsum("ef,bcft->bcet", torch.kron(_BT, _BT), tiles).float().double()                 # fp32 transform results
    U = torch.einsum("ef,ocf->oce", torch.kron(_G, _G), w.reshape(co, C, 9))                        # host side, exact
    Vh, Uh = rnd(V), rnd(U)
    Vl, Ul = rnd(V - Vh), rnd(U - Uh)
    M = torch.einsum("oce,bcet->boet", Uh, Vh) + torch.einsum("oce,bcet->boet", Ul, Vh) + torch.einsum("oce,bcet->boet", Uh, Vl)
    M = M.float().double()                                                                          # fp32 accumulators
    Y = torch.einsum("ye,boet->boyt", torch.kron(_AT, _AT), M)                                      # [B, co, 4, T]
    return F.fold(Y.reshape(B, co * 4, -1), (H, W), kernel_size=2, stride=2)


def make_conv(scheme):
    def conv(x, w, b=None, padding=0):
        x = x.double(); w = w.double()
        if scheme.startswith("wino3") and w.shape[2] == 3 and w.shape[3] == 3 and padding == 1:
            out = winograd3x3(x, w, rbf if "bf16" in scheme else r16)
        elif scheme == "exact":
            out = F.conv2d(x, w, padding=padding)
        else:
            rnd = rbf if scheme.startswith("bf16") else r16
            xh, wh = rnd(x), rnd(w)
            xl, wl = rnd(x - xh), rnd(w - wh)
            out = F.conv2d(xh, wh, padding=padding)
            if scheme in ("fp16x3", "bf16x3", "wino3-fp16x3", "wino3-bf16x3"):
                out = out + F.conv2d(xh, wl, padding=padding) + F.conv2d(xl, wh, padding=padding)
            elif scheme.startswith("mixed"):            # mixed-<fmt>[-noscale]: corrections with 8-bit operands
                fmt = scheme.split("-")[1]
                sc = not scheme.endswith("noscale")
                out = out + F.conv2d(r8(x, fmt, sc), r8(w - wh, fmt, sc), padding=padding) + \
                    F.conv2d(r8(x - xh, fmt, sc), r8(w, fmt, sc), padding=padding)
            elif scheme == "fp16x1":
                pass
            else:
                raise ValueError(scheme)
        if b is not None:
            out = out + b.double().view(1, -1, 1, 1)
        return out.float()                               # fp32 activations between layers, as in the engine
    return conv


def resblock(conv, sd, p, x, k):
    out = F.relu(conv(x, sd[p + ".left.0.weight"], padding=k // 2))
    out = conv(out, sd[p + ".left.2.weight"], padding=k // 2)
    key = p + ".shortcut.0.weight"
    return F.relu(out + (conv(x, sd[key]) if key in sd else x))


def q_net(conv, sd, x, luma):
    ov, k12 = (4, 5) if luma else (2, 3)
    x2 = F.relu(conv(F.pad(x, (0, ov, 0, ov)), sd["conv_q1.weight"], sd["conv_q1.bias"]))
    x3 = resblock(conv, sd, "resblock_q1", x2, k12)
    if luma:
        x3 = F.max_pool2d(x3, 2)
    x4 = F.max_pool2d(resblock(conv, sd, "resblock_q2", x3, k12), 2)
    x5 = resblock(conv, sd, "resblock_q3", x4, 3)
    x6 = torch.cat([x5] + [F.interpolate(F.max_pool2d(x5, s), scale_factor=s) for s in (2, 4, 8)], 1)
    x7 = resblock(conv, sd, "resblock_q4", x6, 3)
    x8 = F.max_pool2d(resblock(conv, sd, "resblock_q5", x7, 3), 2)
    x9 = resblock(conv, sd, "resblock_q6", x8, 3)
    return conv(x9, sd["conv_q2.weight"], sd["conv_q2.bias"], padding=1)


@torch.no_grad()
def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    torch.set_num_threads(os.cpu_count() or 1)
    by, bu, bv = synth.synth_blocks(n, seed=5)
    luma_x = torch.from_numpy(by.astype(np.float32)).unsqueeze(1)
    chroma_x = torch.cat([F.max_pool2d(luma_x, 2), torch.from_numpy(bu.astype(np.float32)).unsqueeze(1),
                          torch.from_numpy(bv.astype(np.float32)).unsqueeze(1)], 1)
    schemes = ["fp16x1", "bf16x3", "fp16x3", "mixed-e4m3", "mixed-e4m3-noscale", "mixed-e5m2", "wino3-fp16x3", "wino3-bf16x3"]
    print("max-abs / mean-abs error of the QT map vs exact (fp64 convs), %d blocks; parity bar 1e-2" % n)
    for comp, x in (("Luma", luma_x), ("Chroma", chroma_x)):
        for qp in (22, 37):
            sd = {k: torch.as_tensor(v) for k, v in load_reference_pkl(os.path.join(ROOT, "trained_models", "%s_Q_%d.pkl" % (comp, qp))).items()}
            ref = q_net(make_conv("exact"), sd, x, comp == "Luma")
            row = []
            for sch in schemes:
                out = q_net(make_conv(sch), sd, x, comp == "Luma")
                e = (out - ref).abs()
                row.append("%s %.1e/%.1e" % (sch, float(e.max()), float(e.mean())))
            print("%s_Q_%d: %s" % (comp, qp, "  ".join(row)), flush=True)


if __name__ == "__main__":
    main()
