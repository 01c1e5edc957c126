"""Drop-in for the reference's ``Model_QBD.py`` (/root/reference/Model_QBD.py:23-253).

Same class names, zero-argument constructors, parameter names/shapes (``state_dict`` contract) and forward
signatures -- ``Luma_Q_Net()(x) -> [B,1,8,8]``, ``Luma_MSBD_Net()(x, qt) -> 3 x [B,2,16,16]`` -- but ``forward``
runs the hand-written sm_100a kernels of libpmp_b200 through the C ABI (``pmp_forward_q`` / ``pmp_forward_msbd``)
instead of cuDNN.  Inference only (no autograd); inputs must live on a CUDA device: there is no CPU fallback.

The modules hold ordinary ``nn.Conv2d`` children purely as parameter containers so that ``load_state_dict`` with the
reference ``.pkl`` files, ``nn.DataParallel`` wrapping and ``.cuda()`` behave exactly as with the reference.  Weights
are packed into the kernels' operand layouts lazily, once per (device, parameter version), in a small LRU per
(device, net kind) so that alternating QPs / DataParallel replicas do not re-pack on every forward.
"""
import ctypes
import itertools
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib
from .netspec import param_spec


class ResidualBlock(nn.Module):
    """Parameter container with the reference layout (Model_QBD.py:23-38): left.0, left.2, shortcut.0."""

    def __init__(self, inchannel, outchannel, kernel_size=3, padding=1, stride=1):
        super().__init__()
        self.left = nn.Sequential(
            nn.Conv2d(inchannel, outchannel, kernel_size=kernel_size, stride=stride, padding=padding, bias=False),
            nn.ReLU(inplace=True),
            nn.Conv2d(outchannel, outchannel, kernel_size=kernel_size, stride=1, padding=padding, bias=False))
        self.shortcut = nn.Sequential()
        if stride != 1 or inchannel != outchannel:
            self.shortcut = nn.Sequential(nn.Conv2d(inchannel, outchannel, kernel_size=1, stride=stride, bias=False))

    def forward(self, x):
        raise RuntimeError("ResidualBlock is fused into the net kernels; call the enclosing net")


def _make_layer(cin, couts, ks):
    layers = []
    for co, k in zip(couts, ks):
        layers.append(ResidualBlock(cin, co, kernel_size=k, padding=k // 2))
        cin = co
    return nn.Sequential(*layers)


_WSET_LRU = 8            # packed weight sets kept per (device, net kind): 4 QPs x (module, DataParallel replica)
_WSET_CACHE = {}         # (device, net kind) -> OrderedDict{fingerprint: wset id}
_SERIAL = itertools.count()      # per-instance serial numbers: id() and data_ptr() are both reused after a module dies


class _PmpNet(nn.Module):
    """Shared plumbing: weight-set cache and the C-ABI call."""
    NET = None

    def __init__(self):
        super().__init__()
        self._serial = next(_SERIAL)   # never reused, unlike id(self): a new net whose tensors land on a dead net's
        #                                addresses (caching allocator) must not hit that net's packed weights
        self._wgen = 0           # bumped by load_state_dict / invalidate_weights(): part of the cache fingerprint
        self._src_fp = None      # set on nn.DataParallel replicas: fingerprint of the module they were replicated from

    def invalidate_weights(self):
        """Force a re-pack of the kernels' operand images on the next forward.  ``load_state_dict`` does this by itself;
        call it after in-place updates through ``param.data`` (those do not bump the parameter's version counter)."""
        self._wgen += 1

    def _load_from_state_dict(self, *args, **kwargs):
        self._wgen += 1
        return super()._load_from_state_dict(*args, **kwargs)

    def _fingerprint(self):
        return (self._serial, self._wgen) + tuple((p.data_ptr(), p._version) for p in self._param_list())

    def _replicate_for_data_parallel(self):
        # runs on the SOURCE module once per replica (torch/nn/parallel/replicate.py): the replica's broadcast weight
        # copies are new tensors every forward, so they are identified by the source's fingerprint (no device sync)
        replica = super()._replicate_for_data_parallel()
        replica._src_fp = self._fingerprint()
        return replica

    def _param_list(self):
        # attribute walk, not named_parameters(): nn.DataParallel replicas on the other GPUs carry their weights as plain
        # tensor attributes (torch/nn/parallel/replicate.py), their _parameters dicts are empty
        out = []
        for name, _ in param_spec(self.NET):
            obj = self
            for part in name.split("."):
                obj = getattr(obj, part)
            out.append(obj)
        return out

    def _weight_set(self, device):
        fp = self._src_fp if getattr(self, "_is_replica", False) and self._src_fp is not None else self._fingerprint()
        cache = _WSET_CACHE.setdefault((device, self.NET), OrderedDict())
        if fp in cache:
            cache.move_to_end(fp)
            return cache[fp]
        h = _lib.Handle.get(device)
        while len(cache) >= _WSET_LRU:
            h.weights_destroy(cache.popitem(last=False)[1])
        cache[fp] = h.weights_create(self.NET, [p.detach() for p in self._param_list()])
        return cache[fp]

    @staticmethod
    def _prep_input(x, channels, size):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise _lib.PmpError("pmp_vvc_tip2023_b200 nets run on a CUDA (sm_100a) device only; got a %s input -- "
                                "there is no CPU fallback" % (x.device if isinstance(x, torch.Tensor) else type(x)))
        if x.dim() != 4 or x.shape[1] != channels or x.shape[2] != size or x.shape[3] != size:
            raise ValueError("expected input [B,%d,%d,%d], got %s" % (channels, size, size, tuple(x.shape)))
        if x.dtype == torch.uint8:
            return x.contiguous(), _lib.IN_U8
        return x.contiguous().float(), _lib.IN_F32


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class _QNet(_PmpNet):
    LUMA = True

    def __init__(self):
        super().__init__()
        cin, k1, k12 = (1, 9, 5) if self.LUMA else (3, 5, 3)
        self.conv_q1 = nn.Conv2d(cin, 32, kernel_size=k1, stride=1, padding=0)
        self.resblock_q1 = ResidualBlock(32, 64, kernel_size=k12, padding=k12 // 2)
        self.resblock_q2 = ResidualBlock(64, 64, kernel_size=k12, padding=k12 // 2)
        self.resblock_q3 = ResidualBlock(64, 32, kernel_size=3, padding=1)
        self.resblock_q4 = ResidualBlock(128, 32, kernel_size=3, padding=1)
        self.resblock_q5 = ResidualBlock(32, 32, kernel_size=3, padding=1)
        self.resblock_q6 = ResidualBlock(32, 8, kernel_size=3, padding=1)
        self.conv_q2 = nn.Conv2d(8, 1, kernel_size=3, stride=1, padding=1)

    @torch.no_grad()
    def forward(self, x):
        x, dt = self._prep_input(x, 1 if self.LUMA else 3, 68 if self.LUMA else 34)
        dev = x.device.index
        B = x.shape[0]
        out = torch.empty((B, 1, 8, 8), dtype=torch.float32, device=x.device)
        if B:
            with torch.cuda.device(dev):
                h = _lib.Handle.get(dev)
                _lib.check(_lib.lib().pmp_forward_q(h.ptr, self._weight_set(dev), x.data_ptr(), dt, B, out.data_ptr(),
                                                    _stream(dev)))
        return out


class _MSBDNet(_PmpNet):
    LUMA = True

    def __init__(self):
        super().__init__()
        cin, kb, ks = (2, 9, 5) if self.LUMA else (4, 5, 3)
        self.conv_b1_1 = nn.Conv2d(cin, 16, kernel_size=kb, stride=1, padding=0)
        self.conv_b1_2 = nn.Conv2d(cin, 8, kernel_size=(ks, kb), stride=1, padding=0)
        self.conv_b1_3 = nn.Conv2d(cin, 8, kernel_size=(kb, ks), stride=1, padding=0)
        self.trunk_M1 = _make_layer(32, [64] * 6, [5, 3, 3, 3, 3, 3])
        self.trunk_M2 = _make_layer(64, [64] * 4, [3] * 4)
        self.trunk_B1 = _make_layer(64, [32, 16, 8], [3] * 3)
        self.trunk_B2 = _make_layer(64, [32, 16, 8], [3] * 3)
        self.trunk_B3 = _make_layer(64, [32, 16, 8], [3] * 3)
        self.conv_B1 = nn.Conv2d(8, 2, kernel_size=3, stride=1, padding=1)
        self.conv_B2 = nn.Conv2d(8, 2, kernel_size=3, stride=1, padding=1)
        self.conv_B3 = nn.Conv2d(8, 2, kernel_size=3, stride=1, padding=1)
        self.trunk_Att1 = _make_layer(3, [32, 64], [3, 3])
        self.trunk_Att2 = _make_layer(3, [32, 64], [3, 3])

    @torch.no_grad()
    def forward(self, x, x1):
        x, dt = self._prep_input(x, 1 if self.LUMA else 3, 68 if self.LUMA else 34)
        if not x1.is_cuda or tuple(x1.shape) != (x.shape[0], 1, 8, 8):
            raise ValueError("expected qt [B,1,8,8] on the input's device, got %s on %s" % (tuple(x1.shape), x1.device))
        qt = x1.contiguous().float()
        dev = x.device.index
        B = x.shape[0]
        outs = [torch.empty((B, 2, 16, 16), dtype=torch.float32, device=x.device) for _ in range(3)]
        if B:
            with torch.cuda.device(dev):
                h = _lib.Handle.get(dev)
                _lib.check(_lib.lib().pmp_forward_msbd(h.ptr, self._weight_set(dev), x.data_ptr(), dt, qt.data_ptr(), B,
                                                       outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(),
                                                       _stream(dev)))
        return outs[0], outs[1], outs[2]


class Luma_Q_Net(_QNet):
    """Model_QBD.py:59-98."""
    NET, LUMA = "Luma_Q", True


class Chroma_Q_Net(_QNet):
    """Model_QBD.py:157-196."""
    NET, LUMA = "Chroma_Q", False


class Luma_MSBD_Net(_MSBDNet):
    """Model_QBD.py:100-155."""
    NET, LUMA = "Luma_MSBD", True


class Chroma_MSBD_Net(_MSBDNet):
    """Model_QBD.py:198-253."""
    NET, LUMA = "Chroma_MSBD", False
