"""A/B timing of the two accumulator schemes of the TC conv kernel (flags bits 8..9: 1 unstacked, 2 stacked)."""
import ctypes, sys
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
for cin, cout, k, hw, b in [(64, 64, 3, 64, 444), (64, 64, 5, 64, 296), (32, 64, 5, 64, 296), (64, 64, 3, 32, 1776), (64, 64, 5, 32, 1184),
                            (64, 32, 3, 32, 1776), (32, 64, 3, 32, 1776), (64, 32, 3, 16, 3552), (128, 32, 3, 16, 3552), (32, 16, 3, 16, 3552), (3, 32, 3, 32, 1776)]:
    for scheme in (1, 2):
        me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
        rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, 3 | (scheme << 8), ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
        fl = 2.0 * b * hw * hw * cin * cout * k * k
        print("cin %3d cout %2d k %d hw %2d B %4d %s: rc %d tc %.3f ms (%.1f TFLOP/s alg) rel err %.2e" % (cin, cout, k, hw, b, "unstacked" if scheme == 1 else "stacked  ", rc, t1.value, fl / max(t1.value, 1e-9) / 1e9, me.value / max(am.value, 1e-9)), flush=True)
