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


# kernel / accumulator-scheme variants (flags: bit 8/9 scheme 1 unstacked 2 stacked, bit 10 CTA-pair kernel, bit 11 force the
# single-CTA kernel, bits 13..15 cap on the activation buffers), odd batches included (the pair kernel's peer CTA then
# recomputes the last image and must drop it), batches large enough for several tiles per cluster (ring wrap-around)
VARIANTS = [((64, 64, 3, 32, 5, 3), 1 << 11), ((64, 64, 3, 32, 5, 3), 1 << 10), ((64, 64, 3, 32, 5, 3), (1 << 10) | (1 << 13)),
            ((64, 64, 5, 64, 3, 1), (1 << 11) | (2 << 8)), ((64, 64, 5, 64, 3, 1), (1 << 11) | (1 << 8)),
            ((64, 32, 3, 16, 7, 1), 1 << 10), ((64, 32, 3, 16, 7, 1), (1 << 11) | (2 << 8)), ((3, 32, 3, 16, 4, 1), 1 << 10),
            ((32, 64, 3, 32, 3, 7), 1 << 10), ((32, 64, 5, 64, 2, 1), 1 << 10), ((64, 64, 5, 64, 5, 3), 1 << 10),
            ((32, 16, 3, 16, 9, 3), 1 << 10), ((32, 16, 3, 16, 9, 3), (1 << 10) | (1 << 13)), ((128, 32, 3, 16, 11, 1), 1 << 10),
            ((64, 32, 1, 16, 7, 0), 1 << 10), ((16, 8, 3, 32, 301, 3), 1 << 10),
            # 3x3 Cout = 64: default stacked scheme on the 4-slot accumulator ring vs the unstacked A/B variant (bit 12),
            # enough tiles per cluster to wrap the slot ring and the weight ring many times
            ((64, 64, 3, 64, 150, 3), 1 << 10), ((64, 64, 3, 64, 150, 3), (1 << 10) | (1 << 12)), ((64, 64, 3, 32, 333, 7), 1 << 10),
            ((64, 64, 3, 16, 75, 7), 1 << 10), ((32, 64, 3, 32, 5, 1), 1 << 11),
            # 5x5 Cout = 64: stacked with single-tap ring stages (default) vs unstacked row stages (bit 12)
            ((64, 64, 5, 64, 150, 3), 1 << 10), ((64, 64, 5, 64, 150, 3), (1 << 10) | (1 << 12)), ((32, 64, 5, 64, 77, 1), 1 << 10),
            ((64, 64, 5, 32, 333, 3), 1 << 10), ((32, 64, 5, 32, 200, 1), 1 << 10), ((64, 64, 5, 64, 3, 1), 1 << 11)]


def _fused(cin2):
    return (1 << 17) | (cin2 << 20)


# 1x1 shortcut conv fused into the conv as extra K groups of a second input (flag bit 17, channels in bits 20..27): every
# ResidualBlock shape of the four nets that has a shortcut (Model_QBD.py:34-38), checked against conv + separate 1x1 conv
VARIANTS += [((64, 64, 5, 64, 5, 1), _fused(32)), ((64, 64, 5, 64, 150, 1), _fused(32)), ((64, 64, 3, 32, 7, 1), _fused(32)),
             ((64, 64, 5, 32, 9, 1), _fused(32)), ((32, 32, 3, 16, 9, 1), _fused(64)), ((32, 32, 3, 16, 301, 1), _fused(128)),
             ((32, 32, 3, 16, 5, 1), _fused(3)), ((32, 32, 3, 32, 6, 1), _fused(3)), ((16, 16, 3, 32, 6, 1), _fused(32)),
             ((16, 16, 3, 16, 7, 1), _fused(32)), ((8, 8, 3, 16, 5, 1), _fused(16)), ((8, 8, 3, 32, 5, 1), _fused(16)),
             ((8, 8, 3, 8, 5, 1), _fused(32)), ((64, 64, 3, 16, 5, 5), _fused(32)), ((64, 64, 3, 32, 201, 5), _fused(32)),
             ((32, 32, 3, 32, 3, 1), _fused(64))]


@pytest.mark.parametrize("cfg,mode", VARIANTS, ids=lambda v: str(v).replace(" ", ""))
def test_tc_conv_variants(cfg, mode):
    cin, cout, k, hw, b, fl = cfg
    h = _lib.Handle.get(0)
    me, am, t1, t2 = (ctypes.c_double() for _ in range(4))
    _lib.check(_lib.lib().pmp_selftest_conv(h.ptr, cin, cout, k, hw, b, fl | mode, ctypes.byref(me), ctypes.byref(am),
                                            ctypes.byref(t1), ctypes.byref(t2)))
    assert me.value <= 2e-5 * max(am.value, 1.0), (me.value, am.value)


# 2x2 max-pool: horizontal half in the conv epilogue (flag bit 18), vertical half by pool2_split; the pooled ResidualBlock
# shapes of the nets (identity residual or fused shortcut)
# a batch of one runs the pair kernel too (the peer CTA recomputes the image and drops it)
VARIANTS += [((64, 64, 3, 32, 1, 3), 0), ((32, 32, 3, 16, 1, 1), _fused(64)), ((64, 64, 5, 64, 1, 3), 1 << 18),
             ((64, 32, 3, 16, 1, 1), 0)]

POOLED = [((64, 64, 5, 64, 5, 1), (1 << 18) | _fused(32)), ((64, 64, 3, 64, 75, 3), 1 << 18), ((64, 64, 5, 32, 9, 3), 1 << 18),
          ((64, 64, 3, 32, 201, 3), 1 << 18), ((32, 32, 3, 16, 7, 3), 1 << 18), ((8, 8, 3, 32, 5, 1), (1 << 18) | _fused(16)),
          ((64, 64, 3, 32, 4, 1), 1 << 18)]


@pytest.mark.parametrize("cfg,mode", POOLED, ids=lambda v: str(v).replace(" ", ""))
def test_tc_conv_fused_hpool(cfg, mode):
    test_tc_conv_variants(cfg, mode)
