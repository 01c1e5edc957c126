"""CTA-pair conv kernel: correctness and timing vs the single-CTA kernel (flags bit 10 = pair, bit 11 = force single)."""
import ctypes, sys
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
nfail = 0
for cin, cout, k, hw, b, fl in [(64, 32, 3, 32, 3, 1), (32, 16, 3, 16, 5, 3), (16, 8, 3, 16, 4, 3), (3, 32, 3, 16, 3, 1), (32, 64, 3, 32, 3, 7), (128, 32, 3, 16, 3, 1), (32, 8, 3, 8, 5, 1),
                                (64, 32, 3, 32, 1776, 1), (64, 32, 3, 16, 3552, 1), (32, 16, 3, 16, 3552, 3), (32, 64, 3, 32, 1776, 7), (64, 64, 3, 64, 2, 1), (64, 64, 3, 64, 5, 3), (32, 64, 5, 64, 3, 1), (64, 64, 5, 32, 4, 3), (64, 64, 3, 32, 7, 3), (32, 64, 1, 64, 3, 0),
                                (64, 64, 3, 64, 444, 3), (64, 64, 5, 64, 296, 3), (32, 64, 5, 64, 296, 1), (64, 64, 3, 32, 1776, 3)]:
    for mode, bit in (("single ", 1 << 11), ("pair   ", 1 << 10)):
        me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
        rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl | bit | (1 << 16), ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
        flp = 2.0 * b * hw * hw * cin * cout * k * k
        rel = me.value / max(am.value, 1e-9)
        ok = rc == 0 and rel < 2e-5
        nfail += (not ok)
        print("cin %3d cout %2d k %d hw %2d B %4d %s: rc %d tc %.3f ms (%.1f TFLOP/s alg) rel err %.2e %s" % (cin, cout, k, hw, b, mode, rc, t1.value, flp / max(t1.value, 1e-9) / 1e9, rel, "OK" if ok else "FAIL " + L.pmp_last_error().decode()), flush=True)
        if rc == -2:
            sys.exit(2)
print("failures:", nfail)
