"""Run the TC-vs-SIMT conv self test over the layer shapes of the four nets (GPU box)."""
import ctypes
import sys

sys.path.insert(0, ".")
from pmp_vvc_tip2023_b200 import _lib

CONFIGS = [  # cin, cout, k, hw, batch, flags (1 relu, 2 residual, 4 mul, 8 bf16)
    (64, 64, 3, 64, 3, 0), (64, 64, 3, 64, 3, 3), (32, 64, 5, 64, 2, 1), (64, 64, 5, 64, 2, 3), (32, 64, 1, 64, 2, 0),
    (64, 64, 3, 32, 5, 3), (64, 64, 5, 32, 3, 3), (32, 64, 3, 32, 2, 7), (3, 32, 3, 32, 2, 1), (64, 32, 3, 32, 2, 1),
    (64, 32, 3, 16, 7, 1), (128, 32, 3, 16, 3, 1), (32, 32, 3, 16, 3, 3), (3, 32, 3, 16, 3, 1), (32, 64, 3, 16, 3, 7),
    (32, 16, 3, 16, 3, 1), (16, 8, 3, 16, 3, 3), (64, 32, 1, 16, 3, 0), (16, 8, 1, 32, 2, 0), (64, 64, 3, 64, 2, 11), (32, 8, 3, 8, 5, 1), (8, 8, 3, 8, 5, 3), (32, 8, 1, 8, 5, 0),
]


def main():
    verbose = 1 << 16
    variants = [int(a) for a in sys.argv[1:]] or [0]
    h = _lib.Handle.get(0)
    L = _lib.lib()
    nfail = 0
    for var in variants:
        for cin, cout, k, hw, b, fl in CONFIGS:
            me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
            rc = L.pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl | verbose | (var << 8), ctypes.byref(me), ctypes.byref(am),
                                     ctypes.byref(t1), ctypes.byref(t2))
            if rc:
                print("variant %d cin %3d cout %2d k %d hw %2d B %d flags %2d: rc %d %s" % (var, cin, cout, k, hw, b, fl, rc,
                                                                                          L.pmp_last_error().decode()))
                nfail += 1
                if rc == -2:
                    return 2           # CUDA error: context is gone
                continue
            rel = me.value / max(am.value, 1e-9)
            ok = rel < (3e-4 if fl & 8 else 2e-5)
            nfail += (not ok)
            print("variant %d cin %3d cout %2d k %d hw %2d B %d flags %2d: max_err %.3e ref_absmax %.3e rel %.2e  tc %.3f ms simt %.3f ms  %s"
                  % (var, cin, cout, k, hw, b, fl, me.value, am.value, rel, t1.value, t2.value, "OK" if ok else "FAIL"))
            sys.stdout.flush()
    print("failures:", nfail)
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
