"""VTM-side binary PartitionMat reader (tools/vtm_reader): the C++ reader fills the arrays
EncAppCfg.cpp:4301-4398 fills, identically from the reference's text format and from the raw int8 form."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from pmp_vvc_tip2023_b200 import partition_io
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _fnv(values):
    h = 1469598103934665603
    for b in np.asarray(values, dtype=np.int8).view(np.uint8).tolist():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


@pytest.fixture(scope="module")
def reader_exe(tmp_path_factory):
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("reader") / "reader_check")
    subprocess.check_call(["g++", "-O2", "-std=c++11", "-o", exe, os.path.join(ROOT, "tools", "vtm_reader", "reader_check.cpp")])
    return exe


@pytest.mark.parametrize("comp", ["Luma", "Chroma"])
def test_reader_text_and_binary_agree(reader_exe, tmp_path, comp):
    src = os.path.join(GOLDEN, "pipeline_%s_QP32_PartitionMat.txt" % comp)          # written by the reference itself
    w, h, nf = cases.PIPE_W, cases.PIPE_H, cases.PIPE_F
    r, c, per = partition_io.values_per_frame(h, w)
    vals = partition_io.text_to_values(src, h, w)
    assert vals.shape == (nf, per)
    # text path of the reader == the values np.loadtxt sees
    base_t = str(tmp_path / "t")
    shutil.copy(src, base_t + ".txt")
    out_t = subprocess.check_output([reader_exe, base_t, str(nf), str(r), str(c)], text=True).split()
    assert out_t == ["txt", _fnv(vals.reshape(-1))]
    # binary path: same arrays
    base_b = str(tmp_path / "b")
    n = partition_io.write_partition_bin(base_b + ".bin", vals, h, w)
    assert n == 32 + vals.size
    back, r2, c2 = partition_io.read_partition_bin(base_b + ".bin")
    assert (r2, c2) == (r, c) and np.array_equal(back, vals)
    out_b = subprocess.check_output([reader_exe, base_b, str(nf), str(r), str(c)], text=True).split()
    assert out_b == ["bin", out_t[1]]


def test_reader_rejects_mismatched_geometry(reader_exe, tmp_path):
    w, h = cases.PIPE_W, cases.PIPE_H
    r, c, per = partition_io.values_per_frame(h, w)
    base = str(tmp_path / "x")
    partition_io.write_partition_bin(base + ".bin", np.zeros((1, per), np.int8), h, w)
    assert subprocess.run([reader_exe, base, "1", str(r + 16), str(c)], capture_output=True).returncode == 1
    assert subprocess.run([reader_exe, base, "2", str(r), str(c)], capture_output=True).returncode == 1      # too few frames
    with pytest.raises(ValueError):
        partition_io.write_partition_bin(base + ".bin", np.zeros((1, per - 1), np.int8), h, w)
