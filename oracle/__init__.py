"""CPU oracle for the partition-map prediction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or the CPU baseline being timed), never as the thing shipped.

Contents
--------
nets_ref.py     plain-PyTorch (CPU, fp32) functional restatement of the four
                Down-Up-CNN forwards of the reference ``Model_QBD.py``.
postproc_ref.py NumPy restatement of ``Metrics.eli_structual_error``.
decode_ref.py   NumPy restatement of ``Map2Partition.py`` (exhaustive map-tree
                search, float32 error sums exactly as NumPy evaluates them).
decode_ref.c    plain-C restatement of the same decode (same enumeration order,
                NumPy's pairwise float32 summation emulated), built by
                ``oracle/Makefile`` into ``oracle/_build/liboracle_decode.so``.
gen_golden.py   imports the UNMODIFIED reference from /root/reference (only
                available in the build container) and freezes input/output
                vectors under ``tests/golden/``.

Pinning status: every restatement here is checked against outputs of the
reference itself (``tests/golden/*.npz`` produced by ``gen_golden.py``).
The one gap: the reference's trained MSBD (``*_BD_*.pkl``) weights are absent
from the mount, so the MSBD nets are pinned with seeded random weights only --
"parity with trained MTT weights unpinned".
"""
