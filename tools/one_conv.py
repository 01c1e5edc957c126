"""One pmp_selftest_conv call: python tools/one_conv.py cin cout k hw batch flags"""
import ctypes, sys
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
cin, cout, k, hw, b, fl = (int(a, 0) for a in sys.argv[1:7])
h = _lib.Handle.get(0); L = _lib.lib()
me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl | (1 << 16), ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
print("rc", rc, L.pmp_last_error().decode() if rc else "", "max_err %.3e ref_absmax %.3e rel %.2e tc %.3f ms" % (me.value, am.value, me.value / max(am.value, 1e-9), t1.value), flush=True)
