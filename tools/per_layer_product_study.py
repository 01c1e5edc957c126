"""CPU accuracy study: which single layers of the trained Q nets tolerate a 2-product scheme (one of the two correction
products of the split-precision conv dropped)?  VERDICT r01 item 2(b).

For every conv layer i of the net, the whole net is evaluated with fp16x3 everywhere except layer i, which drops
`a_lo*w_hi` (activations rounded to fp16 in that layer) or `a_hi*w_lo` (weights rounded to fp16); the max-abs error of the
8x8 QT map against fp64 convolutions is printed per layer, followed by the cumulative error when every layer under a
per-layer budget is switched together.  Pure PyTorch on the CPU (fp64 emulation of the operand rounding).

    python tools/per_layer_product_study.py [n_blocks] [net ...]        e.g. 64 Luma_Q_22
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
from tools.mixed_kind_study import q_net, r16  # noqa: E402


class LayerwiseConv:
    """conv callable for mixed_kind_study.q_net: layer index -> scheme ('x3', 'no_alo', 'no_wlo', 'exact')."""

    def __init__(self, schemes, default="x3"):
        self.schemes, self.default, self.i, self.names = schemes, default, 0, []

    def __call__(self, x, w, b=None, padding=0):
        sch = self.schemes.get(self.i, self.default)
        self.names.append("%dx%d %d->%d @%d" % (w.shape[2], w.shape[3], w.shape[1], w.shape[0], x.shape[-1]))
        self.i += 1
        x = x.double(); w = w.double()
        if sch == "exact":
            out = F.conv2d(x, w, padding=padding)
        else:
            xh, wh = r16(x), r16(w)
            xl, wl = r16(x - xh), r16(w - wh)
            out = F.conv2d(xh, wh, padding=padding)
            if sch != "no_wlo":
                out = out + F.conv2d(xh, wl, padding=padding)
            if sch != "no_alo":
                out = out + F.conv2d(xl, wh, padding=padding)
        if b is not None:
            out = out + b.double().view(1, -1, 1, 1)
        return out.float()


@torch.no_grad()
def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    nets = sys.argv[2:] or ["Luma_Q_22", "Chroma_Q_37"]
    torch.set_num_threads(os.cpu_count() or 1)
    by, bu, bv = synth.synth_blocks(n, seed=5)
    luma_x = torch.from_numpy(by.astype(np.float32)).unsqueeze(1)
    chroma_x = torch.cat([F.max_pool2d(luma_x, 2), torch.from_numpy(bu.astype(np.float32)).unsqueeze(1),
                          torch.from_numpy(bv.astype(np.float32)).unsqueeze(1)], 1)
    for net in nets:
        comp, _, qp = net.split("_")
        luma = comp == "Luma"
        x = luma_x if luma else chroma_x
        sd = {k: torch.as_tensor(v) for k, v in load_reference_pkl(os.path.join(ROOT, "trained_models", net + ".pkl")).items()}
        ref = q_net(LayerwiseConv({}, "exact"), sd, x, luma)
        base = LayerwiseConv({})
        e0 = float((q_net(base, sd, x, luma) - ref).abs().max())
        nl = base.i
        macs = []
        print("%s: %d conv layers, fp16x3 everywhere: max-abs %.2e (parity bar 1e-2)" % (net, nl, e0), flush=True)
        per = {}
        for which in ("no_alo", "no_wlo"):
            for i in range(nl):
                out = q_net(LayerwiseConv({i: which}), sd, x, luma)
                per[(i, which)] = float((out - ref).abs().max())
                print("  layer %2d %-22s %s: max-abs %.2e" % (i, base.names[i], which, per[(i, which)]), flush=True)
        for budget in (2e-4, 5e-4, 1e-3):
            pick = {}
            for i in range(nl):
                best = min(("no_alo", "no_wlo"), key=lambda wch: per[(i, wch)])
                if per[(i, best)] <= budget:
                    pick[i] = best
            out = q_net(LayerwiseConv(pick), sd, x, luma)
            print("  per-layer budget %.0e: %d of %d layers on 2 products (%s) -> whole-net max-abs %.2e" % (
                budget, len(pick), nl, ",".join(str(i) for i in sorted(pick)), float((out - ref).abs().max())), flush=True)


if __name__ == "__main__":
    main()
