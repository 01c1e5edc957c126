"""PartitionMat file formats (host side, no CUDA needed).

Text: one decimal integer per line, per frame ``hor | ver | qt | dire`` (Map2Partition.py:400-412) -- what the patched
VTM-10.0 parses (EncAppCfg.cpp:4301-4398).  Binary (opt-in): the same values as raw int8
behind a 32-byte header; the VTM-side reader is tools/vtm_reader/pmp_partition_reader.h.
"""
import struct

import numpy as np

BIN_MAGIC = b"PMPPART1"


def values_per_frame(height, width):
    """R = 16*(H>>6), C = 16*(W>>6) 4x4 units of the 64-cropped frame (EncAppCfg.cpp:4246-4249)."""
    r, c = 16 * (height >> 6), 16 * (width >> 6)
    return r, c, 2 * r * c + (r // 2) * (c // 2) + 3 * r * c


def bin_header(frames, rows, cols):
    return BIN_MAGIC + struct.pack("<6i", int(frames), int(rows), int(cols), 0, 0, 0)


def write_partition_bin(path, values, height, width):
    """values: int8 array [F, per-frame values] (what PartitionPredictor.predict_frames returns, on the host)."""
    v = np.ascontiguousarray(values, dtype=np.int8)
    r, c, per = values_per_frame(height, width)
    if v.ndim != 2 or v.shape[1] != per:
        raise ValueError("expected [frames, %d] values for %dx%d, got %s" % (per, width, height, v.shape))
    with open(path, "wb") as fp:
        fp.write(bin_header(v.shape[0], r, c))
        fp.write(v.tobytes())
    return 32 + v.size


def read_partition_bin(path):
    with open(path, "rb") as fp:
        head = fp.read(32)
        if head[:8] != BIN_MAGIC:
            raise ValueError("%s is not a PMPPART1 file" % path)
        frames, rows, cols = struct.unpack("<3i", head[8:20])
        per = 2 * rows * cols + (rows // 2) * (cols // 2) + 3 * rows * cols
        v = np.frombuffer(fp.read(frames * per), dtype=np.int8)
    return v.reshape(frames, per), rows, cols


def text_to_values(path, height, width):
    """Parse a reference-format text file into the int8 [F, per-frame] array."""
    _, _, per = values_per_frame(height, width)
    v = np.loadtxt(path, dtype=np.int64)
    if v.size % per:
        raise ValueError("%s: %d values is not a multiple of %d" % (path, v.size, per))
    return v.astype(np.int8).reshape(-1, per)
