"""ctypes binding of libpmp_b200.so (include/pmp_b200.h).  No CPU fallback: every entry point
raises if the library is missing or no B200 is visible."""
import ctypes
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# PMP_B200_LIB: file name of another build of the same library inside this directory (kernel A/B runs in one GPU call)
LIB_PATH = os.path.join(HERE, os.path.basename(os.environ.get("PMP_B200_LIB", "libpmp_b200.so")))

NET_LUMA_Q, NET_LUMA_MSBD, NET_CHROMA_Q, NET_CHROMA_MSBD = 0, 1, 2, 3
NET_IDS = {"Luma_Q": 0, "Luma_MSBD": 1, "Chroma_Q": 2, "Chroma_MSBD": 3}
ENGINE_SIMT, ENGINE_TC = 0, 1
TC_FP16, TC_BF16 = 0, 1
IN_U8, IN_F32 = 0, 1

c_void_p, c_int, c_int64, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_char_p
_P = ctypes.POINTER

# name -> (restype, argtypes); kept in one table so tests can check the ABI against the header
SIGNATURES = {
    "pmp_version": (c_int, []),
    "pmp_last_error": (c_char_p, []),
    "pmp_create": (c_int, [c_int, _P(c_void_p)]),
    "pmp_destroy": (None, [c_void_p]),
    "pmp_set_engine": (c_int, [c_void_p, c_int, c_int]),
    "pmp_get_engine": (c_int, [c_void_p]),
    "pmp_set_near_tol": (c_int, [c_void_p, ctypes.c_float]),
    "pmp_saturation_count": (c_int, [c_void_p, _P(ctypes.c_longlong), c_int]),
    "pmp_launch_count": (ctypes.c_longlong, [c_void_p]),
    "pmp_profile": (c_int, [c_void_p, c_int]),
    "pmp_profile_read": (c_int, [c_void_p, c_int, c_char_p, c_int, _P(ctypes.c_double), _P(ctypes.c_longlong),
                                 _P(ctypes.c_double), _P(ctypes.c_double)]),
    "pmp_weights_create": (c_int, [c_void_p, c_int, _P(c_void_p), _P(c_int64), c_int, _P(c_int)]),
    "pmp_weights_destroy": (c_int, [c_void_p, c_int]),
    "pmp_forward_q": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pmp_forward_msbd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "pmp_predict_maps": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p]),
    "pmp_qt_postprocess": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "pmp_map2partition": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "pmp_map2partition_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, _P(ctypes.c_double), c_void_p,
                                     ctypes.c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pmp_assemble_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                    c_void_p]),
    "pmp_frame_values": (c_int64, [c_int, c_int]),
    "pmp_format_text": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, _P(c_int64), c_void_p]),
    "pmp_cut_blocks": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p]),
    "pmp_run_component": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pmp_selftest_conv": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, _P(ctypes.c_double),
                                  _P(ctypes.c_double), _P(ctypes.c_double), _P(ctypes.c_double)]),
    "pmp_debug_conv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                               c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pmp_debug_stem": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "pmp_debug_tc_stalls": (c_int, [_P(ctypes.c_uint64), c_int]),
}

_lib = None
_lock = threading.Lock()


class PmpError(RuntimeError):
    pass


def lib():
    """Load libpmp_b200.so (raises if it has not been built: there is no fallback path)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PmpError("libpmp_b200.so is not built (%s); run `python -m pmp_vvc_tip2023_b200.build` "
                               "-- there is no CPU/PyTorch fallback" % LIB_PATH)
            L = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)
                fn.restype = res
                fn.argtypes = args
            _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().pmp_last_error()
        raise PmpError("libpmp_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


class Handle:
    """A pmp_handle (engine selection, weight sets, activation arena).  ``Handle.get(device)`` is the per-(process, device)
    default used by the Model_QBD modules and the module-level ops; a PartitionPredictor owns a private one, so that its
    engine choice and arena are its own.  Not thread-safe; one stream at a time per handle."""
    _cache = {}

    def __init__(self, device=0):
        self.device = int(device)
        self._h = c_void_p()
        check(lib().pmp_create(self.device, ctypes.byref(self._h)))
        self._wsets = []

    @classmethod
    def get(cls, device=0):
        device = int(device)
        if device not in cls._cache:
            cls._cache[device] = cls(device)
        return cls._cache[device]

    @property
    def ptr(self):
        return self._h

    def close(self):
        if self._h:
            lib().pmp_destroy(self._h)
            self._h = c_void_p()
        if Handle._cache.get(self.device) is self:
            Handle._cache.pop(self.device, None)

    # ---- engine / counters ------------------------------------------------------------------
    def set_engine(self, engine, tc_dtype=TC_FP16):
        check(lib().pmp_set_engine(self._h, engine, tc_dtype))

    def engine(self):
        return lib().pmp_get_engine(self._h)

    def set_near_tol(self, tol):
        check(lib().pmp_set_near_tol(self._h, float(tol)))

    def saturation_count(self, reset=False):
        """fp16 range-guard events of the TC conv epilogues (non-zero => results clamped: use tc_dtype='bf16')."""
        n = ctypes.c_longlong()
        check(lib().pmp_saturation_count(self._h, ctypes.byref(n), 1 if reset else 0))
        return int(n.value)

    def launch_count(self):
        return int(lib().pmp_launch_count(self._h))

    def profile(self, mode):
        check(lib().pmp_profile(self._h, mode))

    def profile_read(self):
        out = {}
        name = ctypes.create_string_buffer(64)
        for i in range(16):
            ms, n, fl, by = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double(), ctypes.c_double()
            rc = lib().pmp_profile_read(self._h, i, name, 64, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl),
                                        ctypes.byref(by))
            if rc != 0:
                break
            out[name.value.decode()] = {"ms": ms.value, "launches": n.value, "flops": fl.value, "bytes": by.value}
        return out

    # ---- weights ----------------------------------------------------------------------------
    def weights_create(self, net, tensors):
        """tensors: list of contiguous float32 numpy arrays / CPU torch tensors in state_dict order."""
        import numpy as np
        arrs = [np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)
                for t in tensors]
        n = len(arrs)
        ptrs = (c_void_p * n)(*[a.ctypes.data for a in arrs])
        numel = (c_int64 * n)(*[a.size for a in arrs])
        wset = c_int()
        net_id = NET_IDS[net] if isinstance(net, str) else int(net)
        check(lib().pmp_weights_create(self._h, net_id, ptrs, numel, n, ctypes.byref(wset)))
        return wset.value

    def weights_destroy(self, wset):
        check(lib().pmp_weights_destroy(self._h, wset))
