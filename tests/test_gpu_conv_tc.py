"""tcgen05 implicit-GEMM conv (TC engine) against the exact fp32 SIMT conv on identical split-precision inputs, over the
layer shapes of the four nets (pmp_selftest_conv through the C ABI)."""
import ctypes

import pytest

from pmp_vvc_tip2023_b200 import _lib
from tools.selftest_tc import CONFIGS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "cin%d_cout%d_k%d_hw%d_b%d_f%d" % c)
def test_tc_conv_matches_simt(cfg):
    cin, cout, k, hw, b, fl = cfg
    h = _lib.Handle.get(0)
    me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
    _lib.check(_lib.lib().pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl, ctypes.byref(me), ctypes.byref(am),
                                            ctypes.byref(t1), ctypes.byref(t2)))
    # 3-product split precision: fp16 hi/lo ~2^-22 relative, bf16 hi/lo ~2^-16
    assert me.value <= (3e-4 if fl & 8 else 2e-5) * max(am.value, 1.0), (me.value, am.value)
