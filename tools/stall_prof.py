"""Per-role barrier-stall profile of the CTA-pair conv kernel (run with PMP_TC_DBG=64 [+ other knobs])."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
names = ["total", "iss:acc_empty", "iss:act_full", "iss:w_full", "wprod:w_empty", "aprod:act_empty", "epi:acc_full", "-", "items", "epi:total"]
shapes = [(64, 64, 3, 64, 592, 3), (64, 64, 5, 64, 592, 3), (64, 64, 3, 32, 2400, 3), (64, 32, 3, 16, 2400, 1), (32, 16, 3, 16, 2400, 1), (32, 64, 1, 64, 592, 0)]
for cin, cout, k, hw, b, fl in shapes:
    me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
    rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl, ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
    buf = (ctypes.c_uint64 * (148 * 16))()
    L.pmp_debug_tc_stalls(buf, 148 * 16)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 16).astype(np.float64)
    lead, peer = a[0::2], a[1::2]
    tot = lead[:, 0].mean()
    print("dbg %s cin %d cout %d k %d hw %d B %d fl %d: %.3f ms, issuer-0 total %.0f clk, items/cluster %.1f, clk/item %.0f" % (
        os.environ.get("PMP_TC_DBG"), cin, cout, k, hw, b, fl, t1.value, tot, lead[:, 8].mean(), tot / max(lead[:, 8].mean(), 1)))
    for i in (1, 2, 3, 4, 5, 6, 9):
        print("    %-16s leader %5.1f%%  peer %5.1f%%" % (names[i], 100 * lead[:, i].mean() / tot, 100 * peer[:, i].mean() / tot))
    sys.stdout.flush()
