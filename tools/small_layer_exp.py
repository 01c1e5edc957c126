"""Timing experiments on small layers of the pair kernel (PMP_TC_DBG knobs; results are wrong by construction)."""
import ctypes, os, sys
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
for cin, cout, k, hw, b, fl in [(64, 32, 3, 16, 2400, 1), (32, 32, 3, 16, 2400, 3), (32, 16, 3, 16, 2400, 1), (64, 64, 3, 16, 2400, 7), (64, 32, 3, 32, 2400, 1),
                                (64, 64, 3, 32, 2400, 3), (64, 64, 3, 64, 592, 3), (32, 64, 1, 64, 592, 0), (16, 8, 1, 16, 2400, 0)]:
    me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
    rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl, ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
    items = (b + 1) // 2
    print("dbg %s cin %3d cout %2d k %d hw %2d B %4d fl %d: rc %d tc %.3f ms  %.2f us per pair-item-round (74 clusters)" % (os.environ.get("PMP_TC_DBG", "0"), cin, cout, k, hw, b, fl, rc, t1.value, t1.value * 1e3 / (items / 74.0)), flush=True)
