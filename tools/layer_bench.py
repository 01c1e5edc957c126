"""Per-layer timing of the TC conv engine over every layer shape of the four nets at bench batch sizes (GPU box).

Each distinct (cin, cout, k, hw, epilogue) shape runs through pmp_selftest_conv (checked against the exact fp32 conv) and
its time is scaled to one 480-block frame; the table shows where a luma+chroma frame spends its conv time.
usage: python tools/layer_bench.py [batch64 [batch_small [extra_flags]]]
"""
import ctypes
import sys

sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib

R, S, M = 1, 2, 4          # relu, residual, attention product


def rb(ci, co, k, hw, last_flags=R | S):
    out = [(ci, co, k, hw, R)]
    if ci != co:
        out.append((ci, co, 1, hw, 0))
    out.append((co, co, k, hw, last_flags))
    return out


def q_net(luma):
    k, s1 = (5, 64) if luma else (3, 32)
    return (rb(32, 64, k, s1) + rb(64, 64, k, 32) + rb(64, 32, 3, 16) + rb(128, 32, 3, 16) + rb(32, 32, 3, 16) +
            rb(32, 8, 3, 8))


def branch(hw):
    return rb(64, 32, 3, hw) + rb(32, 16, 3, hw) + rb(16, 8, 3, hw)


def att(hw):
    return rb(3, 32, 3, hw) + rb(32, 64, 3, hw, R | S | M)


def msbd_net(luma):
    s1 = 64 if luma else 32
    out = rb(32, 64, 5, s1)
    for _ in range(5):
        out += rb(64, 64, 3, s1)
    for _ in range(4):
        out += rb(64, 64, 3, 32)
    return out + branch(16) + att(16) + branch(16) + att(32) + branch(32)


NETS = {"Luma_Q": q_net(True), "Luma_MSBD": msbd_net(True), "Chroma_Q": q_net(False), "Chroma_MSBD": msbd_net(False)}


def main():
    b64 = int(sys.argv[1]) if len(sys.argv) > 1 else 592
    bsm = int(sys.argv[2]) if len(sys.argv) > 2 else 2400
    extra = int(sys.argv[3], 0) if len(sys.argv) > 3 else 0
    h = _lib.Handle.get(0)
    L = _lib.lib()
    shapes = {}
    for net, layers in NETS.items():
        for l in layers:
            shapes.setdefault(l, {}).setdefault(net, 0)
            shapes[l][net] += 1
    res = {}
    for (cin, cout, k, hw, fl) in sorted(shapes, key=lambda t: (-t[3], -t[0] * t[1] * t[2] * t[2], t[4])):
        b = b64 if hw == 64 else bsm
        me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
        rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl | extra, ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1),
                                 ctypes.byref(t2))
        if rc:
            print("cin %3d cout %2d k %d hw %2d fl %d: rc %d %s" % (cin, cout, k, hw, fl, rc, L.pmp_last_error().decode()))
            if rc == -2:
                return 2
            continue
        flops = 2.0 * hw * hw * cin * cout * k * k
        us480 = t1.value * 1e3 * 480.0 / b
        res[(cin, cout, k, hw, fl)] = us480
        rel = me.value / max(am.value, 1e-9)
        uses = " ".join("%s x%d" % (n, c) for n, c in shapes[(cin, cout, k, hw, fl)].items())
        print("cin %3d cout %2d k %d hw %2d fl %d B %4d: %8.3f ms  %7.1f us/480blk  %6.1f TFLOP/s alg  rel %.1e %s | %s"
              % (cin, cout, k, hw, fl, b, t1.value, us480, flops * b / max(t1.value, 1e-9) / 1e9, rel,
                 "OK" if rel < 2e-5 else "FAIL", uses), flush=True)
    total = 0.0
    for net, layers in NETS.items():
        t = sum(res.get(l, 0.0) for l in layers)
        fl = sum(2.0 * l[3] * l[3] * l[0] * l[1] * l[2] * l[2] for l in layers) * 480
        total += t
        print("%-12s %3d TC convs (stem excluded): %8.1f us per 480 blocks, %6.1f TFLOP/s alg" % (net, len(layers), t, fl / t / 1e6))
    print("all four nets: %.1f us per 480 blocks -> %.0f CTU/s from these convs alone" % (total, 120.0 / (total * 1e-6)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
