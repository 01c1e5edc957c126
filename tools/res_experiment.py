"""Residual-epilogue experiment: the same conv with / without an identity residual at HBM-sized and L2-sized batches,
and (with PMP_TC_DBG=64) per-role stall counters of the small residual layers."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
names = ["total", "iss:acc_empty", "iss:act_full", "iss:w_full", "wprod:w_empty", "aprod:act_empty", "epi:acc_full", "-", "items", "epi:total"]
prof = (int(os.environ.get("PMP_TC_DBG", "0")) & 64) != 0
shapes = [(64, 64, 3, 64, 592), (64, 64, 3, 64, 74), (64, 64, 3, 64, 36), (64, 64, 3, 32, 2400), (64, 64, 3, 32, 148),
          (32, 32, 3, 32, 2400), (16, 16, 3, 32, 2400), (8, 8, 3, 32, 2400), (32, 32, 3, 16, 2400), (16, 16, 3, 16, 2400),
          (64, 32, 3, 16, 2400), (32, 64, 3, 16, 2400)]
for cin, cout, k, hw, b in shapes:
    for fl in (1, 3, 7):
        ts = []
        for rep in range(3):
            me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
            rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl, ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
            if rc:
                print("rc", rc, L.pmp_last_error().decode()); break
            ts.append(t1.value)
        print("cin %3d cout %3d k %d hw %2d B %4d fl %d: %s ms  (%.1f us/480blk)" % (cin, cout, k, hw, b, fl, " ".join("%.4f" % t for t in ts), min(ts) * 1e3 * 480 / b), flush=True)
        if prof:
            buf = (ctypes.c_uint64 * (148 * 16))()
            L.pmp_debug_tc_stalls(buf, 148 * 16)
            a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 16).astype(np.float64)
            lead, peer = a[0::2], a[1::2]
            tot = lead[:, 0].mean()
            print("    issuer-0 total %.0f clk, items/cluster %.1f, clk/item %.0f | " % (tot, lead[:, 8].mean(), tot / max(lead[:, 8].mean(), 1)) +
                  "  ".join("%s %.1f%%" % (names[i], 100 * lead[:, i].mean() / tot) for i in (1, 2, 3, 4, 5, 6, 9)), flush=True)
