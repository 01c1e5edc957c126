"""A/B of the stacked (default) vs unstacked (flag bit 12) accumulator scheme on the 3x3 Cout = 64 layers, pair kernel."""
import ctypes, sys
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
for cin, cout, k, hw, b, fl in [(64, 64, 5, 64, 592, 3), (32, 64, 5, 64, 592, 1), (64, 64, 5, 32, 2400, 1), (64, 64, 5, 32, 2400, 3), (32, 64, 5, 32, 2400, 1), (64, 64, 3, 64, 592, 1), (64, 64, 3, 64, 592, 3), (64, 64, 3, 32, 2400, 1), (64, 64, 3, 32, 2400, 3), (64, 64, 3, 32, 2400, 7), (32, 64, 3, 32, 2400, 1), (64, 64, 3, 16, 2400, 7), (32, 64, 3, 16, 2400, 1)]:
    for name, bit in (("stacked  ", 0), ("unstacked", 1 << 12)):
        me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
        rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl | bit, ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
        flp = 2.0 * b * hw * hw * cin * cout * k * k
        print("cin %3d cout %2d k %d hw %2d B %4d fl %d %s: rc %d %.3f ms  %6.1f us/480  %6.1f TFLOP/s alg  rel %.1e" % (cin, cout, k, hw, b, fl, name, rc, t1.value, t1.value * 1e3 * 480 / b, flp / max(t1.value, 1e-9) / 1e9, me.value / max(am.value, 1e-9)), flush=True)
