import ctypes, sys
sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib
h = _lib.Handle.get(0); L = _lib.lib()
cin, cout, k, hw, b, fl = 64, 64, 3, 64, 444, 1
me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl | (1 << 10), ctypes.byref(me), ctypes.byref(am), ctypes.byref(t1), ctypes.byref(t2))
print(rc, t1.value)
