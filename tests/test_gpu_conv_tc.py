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


# Accumulator-slot reuse across issuer warps (round 2): with 4 slots and 3 M-tiles per item (stacked Cout = 64) consecutive
# uses of a slot belong to different issuer warps; a slow epilogue (residual + attention operand, little MMA work per item)
# lets one issuer run two uses ahead of another, where a bare parity wait aliases.  These shapes hung before the issue
# tickets of pair_issuer; several items per cluster are needed (batch >= 1200).
@pytest.mark.timeout(180)
@pytest.mark.parametrize("cfg", [(64, 64, 3, 16, 1200, 6), (64, 64, 3, 16, 2400, 7), (32, 64, 3, 16, 2400, 3), (32, 64, 3, 16, 2400, 7),
                                 (64, 64, 3, 8, 2400, 7), (32, 64, 5, 16, 2401, 7)],
                         ids=lambda c: "cin%d_cout%d_k%d_hw%d_b%d_f%d" % c)
def test_tc_conv_slow_epilogue_slot_reuse(cfg):
    test_tc_conv_matches_simt(cfg)


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


# ---------------------------------------------------------------------------------------------------------------------
# Independent reference: the same kernels against torch's CPU F.conv2d in float64 (pmp_debug_conv / pmp_debug_stem feed
# caller-supplied fp32 data through the tcgen05 path), one case per kernel-variant class -- a layout or padding bug shared
# by conv_tc_pair_kernel and this library's own SIMT conv (the self test above) cannot hide here.
# ---------------------------------------------------------------------------------------------------------------------
import zlib

import numpy as np
import torch
import torch.nn.functional as F

from pmp_vvc_tip2023_b200 import netspec, synth

RELU, BF16, POOL = 1, 8, 1 << 18


def _debug_conv(x, w, res=None, mul=None, x2=None, wsc=None, flags=0):
    h = _lib.Handle.get(0)
    B, cin, hw, _ = x.shape
    cout, k = w.shape[0], w.shape[2]
    ho = hw // 2 if flags & POOL else hw
    out = torch.empty((B, cout, ho, ho), dtype=torch.float32, device="cuda")
    dev = [None if t is None else t.contiguous().cuda() for t in (x, res, mul, x2)]
    wn = np.ascontiguousarray(w.numpy(), np.float32)
    wscn = None if wsc is None else np.ascontiguousarray(wsc.numpy().reshape(cout, -1), np.float32)
    ptr = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().pmp_debug_conv(h.ptr, ptr(dev[0]), wn.ctypes.data, ptr(dev[1]), ptr(dev[2]), ptr(dev[3]),
                                         None if wscn is None else wscn.ctypes.data, cin, cout, k, hw, B,
                                         0 if x2 is None else x2.shape[1], flags, ptr(out), None))
    return out.cpu()


def _torch_ref(x, w, res=None, mul=None, x2=None, wsc=None, flags=0):
    o = F.conv2d(x.double(), w.double(), padding=w.shape[2] // 2)
    if x2 is not None:
        o = o + F.conv2d(x2.double(), wsc.double())
    if res is not None:
        o = o + res.double()
    if flags & RELU:
        o = F.relu(o)
    if flags & POOL:
        o = F.max_pool2d(o, 2)
    if mul is not None:
        o = o * mul.double()
    return o


# (class, cin, cout, k, hw, B, flags, residual, attention product, fused-shortcut channels)
TORCH_CASES = [
    ("stacked_cout64_3x3", 64, 64, 3, 64, 3, RELU, False, False, 0),
    ("stacked_cout64_5x5_single_tap_stages", 64, 64, 5, 64, 2, RELU, False, False, 0),
    ("identity_residual", 64, 64, 3, 32, 5, RELU, True, False, 0),
    ("fused_shortcut_5x5", 64, 64, 5, 64, 3, RELU, False, False, 32),
    ("fused_shortcut_cin128", 32, 32, 3, 16, 7, RELU, False, False, 128),
    ("fused_shortcut_cin3_padded", 32, 32, 3, 32, 4, RELU, False, False, 3),
    ("fused_shortcut_attention_product", 64, 64, 3, 32, 3, RELU, False, True, 32),
    ("half_pool_residual", 64, 64, 3, 64, 3, RELU | POOL, True, False, 0),
    ("half_pool_fused_shortcut", 64, 64, 5, 64, 2, RELU | POOL, False, False, 32),
    ("half_pool_small", 32, 32, 3, 16, 9, RELU | POOL, True, False, 0),
    ("cout32_wide_k", 128, 32, 3, 16, 5, RELU, False, False, 0),
    ("cout16", 32, 16, 3, 32, 6, RELU, False, False, 0),
    ("cout8_8x8", 32, 8, 3, 8, 11, RELU, False, False, 0),
    ("cin3_first_att_conv", 3, 32, 3, 16, 4, RELU, False, False, 0),
    ("conv1x1_shortcut_unfused", 32, 64, 1, 32, 3, 0, False, False, 0),
    ("batch_of_one", 64, 64, 3, 32, 1, RELU, True, False, 0),
    ("many_tiles_ring_wrap", 64, 64, 3, 32, 301, RELU, True, False, 0),
    ("bf16_split", 64, 64, 3, 32, 3, RELU | BF16, True, False, 0),
]


@pytest.mark.parametrize("case", TORCH_CASES, ids=lambda c: c[0])
def test_tc_conv_matches_torch_conv2d(case):
    name, cin, cout, k, hw, B, flags, use_res, use_mul, cin2 = case
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
    nd = min(B, 6)                                  # distinct images; larger batches repeat them
    rep = lambda t: t.repeat((B + nd - 1) // nd, 1, 1, 1)[:B].contiguous()
    x = rep(F.relu(torch.randn((nd, cin, hw, hw), generator=g)) * 40.0)
    w = torch.randn((cout, cin, k, k), generator=g) * (3.0 / (cin * k * k)) ** 0.5
    ho = hw // 2 if flags & POOL else hw
    res = rep(torch.randn((nd, cout, hw, hw), generator=g) * 10.0) if use_res else None
    mul = rep(torch.randn((nd, cout, ho, ho), generator=g) * 1.5) if use_mul else None
    x2 = rep(F.relu(torch.randn((nd, cin2, hw, hw), generator=g)) * 30.0) if cin2 else None
    wsc = torch.randn((cout, cin2, 1, 1), generator=g) * (3.0 / cin2) ** 0.5 if cin2 else None
    got = _debug_conv(x, w, res, mul, x2, wsc, flags)
    want = _torch_ref(x, w, res, mul, x2, wsc, flags)
    err = float((got.double() - want).abs().max())
    scale = max(float(want.abs().max()), 1.0)
    # 3-product split precision: ~2^-22 relative (fp16 hi/lo), ~2^-16 (bf16 hi/lo)
    assert err <= (3e-4 if flags & BF16 else 2e-5) * scale, (name, err, scale)


@pytest.mark.parametrize("net", ["Luma_Q", "Chroma_Q", "Luma_MSBD", "Chroma_MSBD"])
def test_tc_stem_mode_matches_torch_conv2d(net):
    """First-layer convs in the pair kernel's stem mode (K chunks assembled by 5-D TMA from pre-shifted copies, the three
    MSBD stems merged into one conv) against torch: conv_q1 on padding_rb(x) (Model_QBD.py:79-80), or conv_b1_1..3 on
    cat[x, pad_lu(up(qt))] with their asymmetric zero pads (:104-106,:130-135)."""
    luma, msbd = net.startswith("Luma"), net.endswith("MSBD")
    S0, S1, ov, up = (68, 64, 4, 8) if luma else (34, 32, 2, 4)
    sd = {k: torch.from_numpy(v) for k, v in synth.seeded_state_dict(net, 77).items()}
    g = torch.Generator().manual_seed(5)
    B = 7                                           # odd: the pair kernel's peer CTA recomputes and drops the last image
    x = torch.randint(0, 256, (B, 1 if luma else 3, S0, S0), generator=g).float()
    qt = (torch.randn((B, 1, 8, 8), generator=g) * 1.1 + 1.3) if msbd else None
    h = _lib.Handle.get(0)
    h.set_engine(_lib.ENGINE_TC)
    wset = h.weights_create(net, [sd[k] for k, _ in netspec.param_spec(net)])
    out = torch.empty((B, 32, S1, S1), dtype=torch.float32, device="cuda")
    xc, qc = x.cuda(), None if qt is None else qt.cuda()
    try:
        _lib.check(_lib.lib().pmp_debug_stem(h.ptr, wset, xc.data_ptr(), _lib.IN_F32, None if qc is None else qc.data_ptr(), B,
                                             out.data_ptr(), None))
        torch.cuda.synchronize()
    finally:
        h.weights_destroy(wset)
    xd = x.double()
    if not msbd:
        want = F.relu(F.conv2d(F.pad(xd, (0, ov, 0, ov)), sd["conv_q1.weight"].double(), sd["conv_q1.bias"].double()))
    else:
        x2 = torch.cat([xd, F.pad(F.interpolate(qt.double(), scale_factor=up), (ov, 0, ov, 0))], 1)
        want = torch.cat([F.relu(F.conv2d(F.pad(x2, pad), sd["conv_b1_%d.weight" % i].double(), sd["conv_b1_%d.bias" % i].double()))
                          for i, pad in ((1, (0, ov, 0, ov)), (2, (0, ov, 0, 0)), (3, (0, 0, 0, ov)))], 1)
    err = float((out.cpu().double() - want).abs().max())
    assert err <= 2e-5 * max(float(want.abs().max()), 1.0), (net, err)
