"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/pmp_b200.h declares, fails
loudly without a GPU, and the host-side mirror of the reference interface is consistent.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from pmp_vvc_tip2023_b200 import _lib, build, netspec, synth
from pmp_vvc_tip2023_b200 import Inference_QBD, Model_QBD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_GPU = not torch.cuda.is_available()


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def _declared():
    src = open(os.path.join(ROOT, "include", "pmp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pmp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libpmp_b200.so does not export %s" % n
        assert n in _lib.SIGNATURES, "%s has no ctypes signature" % n
    assert sorted(_lib.SIGNATURES) == names
    assert lib.pmp_version() == 200


def test_frame_values(lib):
    assert lib.pmp_frame_values(3, 6) == 24192            # 416x240
    assert lib.pmp_frame_values(16, 30) == 645120         # 1920x1080


@pytest.mark.skipif(not NO_GPU, reason="checks the no-device failure mode")
def test_create_fails_loudly_without_gpu(lib):
    h = ctypes.c_void_p()
    rc = lib.pmp_create(0, ctypes.byref(h))
    assert rc == -3 and not h.value
    assert b"no CPU fallback" in lib.pmp_last_error()
    with pytest.raises(_lib.PmpError):
        _lib.Handle(0)


def test_modules_refuse_cpu_inputs():
    net = Model_QBD.Luma_Q_Net()
    with pytest.raises(_lib.PmpError):
        net(torch.zeros(1, 1, 68, 68))


@pytest.mark.parametrize("cls", ["Luma_Q_Net", "Luma_MSBD_Net", "Chroma_Q_Net", "Chroma_MSBD_Net"])
def test_state_dict_contract(cls):
    m = getattr(Model_QBD, cls)()
    sd = m.state_dict()
    spec = netspec.param_spec(m.NET)
    assert [k for k, _ in spec] == list(sd.keys())
    assert all(tuple(sd[k].shape) == tuple(s) for k, s in spec)
    # reference .pkl files carry a 'module.' prefix and load through load_pretrain_model
    if cls.endswith("_Q_Net"):
        comp = cls.split("_")[0]
        Inference_QBD.load_pretrain_model(m, os.path.join(ROOT, "trained_models", "%s_Q_32.pkl" % comp))
        raw = torch.load(os.path.join(ROOT, "trained_models", "%s_Q_32.pkl" % comp), map_location="cpu",
                         weights_only=False)
        assert torch.equal(m.state_dict()["conv_q1.weight"], raw["module.conv_q1.weight"])
    # shape-mismatched / unknown keys are skipped silently (Inference_QBD.py:40-43)
    bad = {"module.conv_q1.weight": torch.zeros(3, 3), "module.nope": torch.zeros(1)}
    Inference_QBD.load_pretrain_model(m, bad)


def test_param_counts():
    assert netspec.param_count("Luma_Q") == 464585 and netspec.param_count("Chroma_Q") == 235017
    assert netspec.param_count("Luma_MSBD") == 1075670 and netspec.param_count("Chroma_MSBD") == 1074198


def test_sequence_info_and_yuv_reader(tmp_path):
    info = tmp_path / "seqs.txt"
    info.write_text("A,A_192x128_30.yuv,192,128,61,30\nB,B_416x240_8bit.yuv,416,240,5,30\n#end!!!!\nX,x,1,1,1,1\n")
    names, paths, w, h, nf, sub, blocks = Inference_QBD.load_sequences_info(str(info), ss_ratio=30)
    assert list(names) == ["A", "B"] and sub == [3, 1] and blocks == [3 * 2 * 3, 6 * 3 * 1]
    y, u, v = synth.synth_yuv420(192, 128, 4, seed=3)
    f = tmp_path / "a.yuv"
    with open(f, "wb") as fp:
        for i in range(4):
            fp.write(y[i].tobytes()); fp.write(u[i].tobytes()); fp.write(v[i].tobytes())
    yy, uu, vv = Inference_QBD.import_yuv420(str(f), 192, 128, 4, SubSampleRatio=2, is10bit=True)
    assert np.array_equal(yy, y[::2]) and np.array_equal(uu, u[::2]) and np.array_equal(vv, v[::2])
