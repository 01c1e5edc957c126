"""Timing experiments on the TC conv kernel (debug modes give wrong results on purpose)."""
import ctypes, sys
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
for cin, cout, k, hw, b in [(64, 64, 3, 64, 444), (64, 64, 5, 64, 296), (64, 64, 3, 32, 1776), (64, 32, 3, 16, 3552)]:
    for dbg in (0, 1, 2, 3):
        me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
        rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, 1 | (dbg << 8), ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
        fl = 2.0 * b * hw * hw * cin * cout * k * k
        print("cin %d cout %d k %d hw %d B %d dbg %d: rc %d tc %.3f ms (%.1f TFLOP/s alg) err %.2e" % (cin, cout, k, hw, b, dbg, rc, t1.value, fl / t1.value / 1e9, me.value), flush=True)
