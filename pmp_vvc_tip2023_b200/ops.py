"""Torch-tensor front end of the C ABI (include/pmp_b200.h).  torch is used for device memory and streams only;
every op below is one call into libpmp_b200 and raises if the tensors are not on a CUDA device."""
import ctypes

import torch

from . import _lib


def _dev(t):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.PmpError("expected a CUDA tensor (no CPU fallback), got %s" %
                            (t.device if isinstance(t, torch.Tensor) else type(t)))
    return t.device.index


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _h(handle, dev):
    """The pmp_handle to run on: the caller's own (PartitionPredictor) or the per-device default."""
    if handle is None:
        return _lib.Handle.get(dev).ptr
    if handle.device != dev:
        raise _lib.PmpError("handle is bound to cuda:%d, tensors live on cuda:%d" % (handle.device, dev))
    return handle.ptr


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def qt_postprocess(qt, want_f32=True, want_u8=True, handle=None):
    """Metrics.eli_structual_error (Metrics.py:612-637).  qt [N,1,8,8] f32 -> ([N,1,8,8] f32, [N,64] u8)."""
    dev = _dev(qt)
    qt = qt.contiguous().float()
    n = qt.shape[0]
    of = torch.empty((n, 1, 8, 8), dtype=torch.float32, device=qt.device) if want_f32 else None
    ou = torch.empty((n, 64), dtype=torch.uint8, device=qt.device) if want_u8 else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pmp_qt_postprocess(_h(handle, dev), _ptr(qt), n, _ptr(of), _ptr(ou), _stream(dev)))
    return of, ou


DEFAULT_LAMB = (0.7, 0.7, 1.5, 0.3, 0.7)          # Map2Partition.py:100

FLAG_NEAR_TIE, FLAG_NEAR_BT, FLAG_NEAR_DIRE, FLAG_NEAR_QT = 1, 2, 4, 8
FLAG_NEAR_THRESHOLD = FLAG_NEAR_BT | FLAG_NEAR_DIRE | FLAG_NEAR_QT
FLAG_NEAR_THRESHOLD_TIGHT = FLAG_NEAR_THRESHOLD << 3        # the same three tests at near_tol / 100


def flag_counts(flags):
    """{near_tie_blocks, near_threshold_blocks, near_threshold_blocks_tight, blocks} of a flags tensor
    (pmp_map2partition flags, include/pmp_b200.h)."""
    f = flags.to(torch.int64)
    return {"blocks": int(f.numel()), "near_tie_blocks": int((f & FLAG_NEAR_TIE).ne(0).sum()),
            "near_threshold_blocks": int((f & FLAG_NEAR_THRESHOLD).ne(0).sum()),
            "near_threshold_blocks_tight": int((f & FLAG_NEAR_THRESHOLD_TIGHT).ne(0).sum())}


def map2partition(qt_u8, bt, dire, chroma_factor, want_flags=True, lamb=None, qt_raw=None, near_tol=None, handle=None):
    """Map2Partition.map_to_parititon over a batch (Map2Partition.py:98-373).

    qt_u8 [N,64] uint8 (0..3), bt/dire [N,3,16,16] f32 -> hor,ver [N,16,16] u8, dire_out [N,3,16,16] i8, flags [N] i32.
    lamb: (lamb1..lamb5) of the Map_to_Partition constructor (:100), None = reference defaults; qt_raw [N,1,8,8] f32 and
    near_tol feed the near-threshold bits of `flags`."""
    dev = _dev(qt_u8)
    n = qt_u8.shape[0]
    qt_u8 = qt_u8.contiguous().view(n, 64)
    bt = bt.contiguous().float()
    dire = dire.contiguous().float()
    assert qt_u8.dtype == torch.uint8 and tuple(bt.shape) == (n, 3, 16, 16) and tuple(dire.shape) == (n, 3, 16, 16)
    hor = torch.empty((n, 16, 16), dtype=torch.uint8, device=bt.device)
    ver = torch.empty((n, 16, 16), dtype=torch.uint8, device=bt.device)
    dout = torch.empty((n, 3, 16, 16), dtype=torch.int8, device=bt.device)
    flags = torch.zeros((n,), dtype=torch.int32, device=bt.device) if want_flags else None
    with torch.cuda.device(dev):
        if lamb is None and qt_raw is None and near_tol is None:
            _lib.check(_lib.lib().pmp_map2partition(_h(handle, dev), _ptr(qt_u8), _ptr(bt), _ptr(dire), n,
                                                    int(chroma_factor), _ptr(hor), _ptr(ver), _ptr(dout), _ptr(flags),
                                                    _stream(dev)))
        else:
            lam = None
            if lamb is not None:
                if len(lamb) != 5:
                    raise ValueError("lamb must hold lamb1..lamb5")
                lam = (ctypes.c_double * 5)(*[float(x) for x in lamb])
            if qt_raw is not None:
                qt_raw = qt_raw.contiguous().float()
                assert qt_raw.numel() == n * 64 and qt_raw.device == bt.device
            _lib.check(_lib.lib().pmp_map2partition_ex(_h(handle, dev), _ptr(qt_u8), _ptr(bt), _ptr(dire), n,
                                                       int(chroma_factor), lam, _ptr(qt_raw),
                                                       1e-2 if near_tol is None else float(near_tol), _ptr(hor), _ptr(ver),
                                                       _ptr(dout), _ptr(flags), _stream(dev)))
    return hor, ver, dout, flags


def frame_values(bh, bw):
    return int(_lib.lib().pmp_frame_values(bh, bw))


def assemble_frames(hor, ver, qt_u8, dire, frames, bh, bw, handle=None):
    """Scatter + per-frame vector order of get_sequence_partition_for_VTM (Map2Partition.py:389-412) -> int8 [F, per]."""
    dev = _dev(hor)
    n = frames * bh * bw
    assert hor.shape[0] == n and ver.shape[0] == n and qt_u8.shape[0] == n and dire.shape[0] == n
    out = torch.empty((frames, frame_values(bh, bw)), dtype=torch.int8, device=hor.device)
    # contiguous copies are bound to names that outlive the launch (a temporary's block could be reused by the next one)
    hor, ver, qt_u8, dire = hor.contiguous(), ver.contiguous(), qt_u8.contiguous(), dire.contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pmp_assemble_frames(_h(handle, dev), _ptr(hor), _ptr(ver), _ptr(qt_u8), _ptr(dire),
                                                  frames, bh, bw, _ptr(out), _stream(dev)))
    return out


def format_text(values, handle=None):
    """The text body of a PartitionMat file (Map2Partition.py:405-412): ``str(v) + '\\n'`` per value -> uint8 tensor."""
    dev = _dev(values)
    v = values.contiguous().view(-1)
    assert v.dtype == torch.int8
    n = v.numel()
    text = torch.empty((3 * n,), dtype=torch.uint8, device=v.device)
    nbytes = ctypes.c_int64(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pmp_format_text(_h(handle, dev), _ptr(v), n, _ptr(text), ctypes.byref(nbytes),
                                              _stream(dev)))
    return text[:nbytes.value]


def cut_blocks(y, u, v, want_luma=True, want_chroma=True, handle=None):
    """Inference_QBD.output_block_yuv + chroma input assembly (Inference_QBD.py:104-149,:194-200) on device.

    y [F,H,W], u/v [F,H/2,W/2]; uint8 (8-bit) or int16/uint16 (10-bit, reduced with round-half-even(y/4)).
    Returns (luma_blocks [N,1,68,68] u8, chroma_blocks [N,3,34,34] u8)."""
    dev = _dev(y)
    f, hgt, wid = y.shape
    sb = y.element_size()
    assert sb in (1, 2) and u.element_size() == sb and v.element_size() == sb
    n = f * (hgt // 64) * (wid // 64)
    lb = torch.empty((n, 1, 68, 68), dtype=torch.uint8, device=y.device) if want_luma else None
    cb = torch.empty((n, 3, 34, 34), dtype=torch.uint8, device=y.device) if want_chroma else None
    y, u, v = y.contiguous(), u.contiguous(), v.contiguous()     # named: they must outlive the launch
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pmp_cut_blocks(_h(handle, dev), _ptr(y), _ptr(u), _ptr(v), sb, f, wid, hgt,
                                             _ptr(lb), _ptr(cb), _stream(dev)))
    return lb, cb


def predict_maps(wset_q, wset_msbd, blocks, handle=None):
    """One batch of Metrics.inference_pre_QBD (Metrics.py:387-419) -> qt [B,1,8,8], bt [B,3,16,16], dire [B,3,16,16]."""
    dev = _dev(blocks)
    blocks = blocks.contiguous()
    dt = _lib.IN_U8 if blocks.dtype == torch.uint8 else _lib.IN_F32
    if dt == _lib.IN_F32:
        blocks = blocks.float()
    b = blocks.shape[0]
    qt = torch.empty((b, 1, 8, 8), dtype=torch.float32, device=blocks.device)
    bt = torch.empty((b, 3, 16, 16), dtype=torch.float32, device=blocks.device)
    dire = torch.empty((b, 3, 16, 16), dtype=torch.float32, device=blocks.device)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pmp_predict_maps(_h(handle, dev), wset_q, wset_msbd, _ptr(blocks), dt, b, _ptr(qt),
                                               _ptr(bt), _ptr(dire), _stream(dev)))
    return qt, bt, dire


def run_component(wset_q, wset_msbd, luma, blocks, frames, bh, bw, chunk=1024, want_maps=False, want_flags=False,
                  handle=None):
    """predict_maps -> qt_postprocess -> map2partition -> assemble_frames for one component.

    blocks: uint8 [F*bh*bw,1,68,68] (luma) or [F*bh*bw,3,34,34] (chroma).  Returns int8 [F, per]; with want_maps
    (out, qt, bt, dire, flags); with want_flags alone (out, flags)."""
    dev = _dev(blocks)
    assert blocks.dtype == torch.uint8
    blocks = blocks.contiguous()
    n = frames * bh * bw
    assert blocks.shape[0] == n
    out = torch.empty((frames, frame_values(bh, bw)), dtype=torch.int8, device=blocks.device)
    qt = bt = dire = flags = None
    if want_maps:
        qt = torch.empty((n, 1, 8, 8), dtype=torch.float32, device=blocks.device)
        bt = torch.empty((n, 3, 16, 16), dtype=torch.float32, device=blocks.device)
        dire = torch.empty((n, 3, 16, 16), dtype=torch.float32, device=blocks.device)
    if want_maps or want_flags:
        flags = torch.zeros((n,), dtype=torch.int32, device=blocks.device)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().pmp_run_component(_h(handle, dev), wset_q, wset_msbd, 1 if luma else 0,
                                                _ptr(blocks), frames, bh, bw, int(chunk), _ptr(out), _ptr(qt), _ptr(bt),
                                                _ptr(dire), _ptr(flags), _stream(dev)))
    if want_maps:
        return out, qt, bt, dire, flags
    return (out, flags) if want_flags else out
