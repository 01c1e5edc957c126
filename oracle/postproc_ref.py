"""NumPy restatement of the QT-map post-process (TEST ORACLE ONLY).

Follows /root/reference/Metrics.py:
  check_square_unity  :612-628
  eli_structual_error :630-637

Pinned against the reference itself by tests/golden/postproc.npz.
"""
import numpy as np


def square_unity(m):
    """Metrics.py:612-628 on one 4x4 integer-valued map (returns a new array)."""
    m = m.copy()
    n0 = int(np.count_nonzero(m == 0))
    if n0 <= 12:
        m[m == 0] = 1
        for i in (0, 2):
            for j in (0, 2):
                sub = m[i:i + 2, j:j + 2]          # view: edits land in m
                s = sub.sum()
                if 5 <= s <= 10:
                    if np.count_nonzero(sub == 1) < 3:
                        sub[sub == 1] = 2
                    else:
                        sub[:, :] = 1
    elif n0 < 16:
        m[:, :] = 0
    return m


def eli_structural_error(qt):
    """Metrics.py:630-637.  qt: [N,1,8,8] float32 -> [N,1,8,8] float32 holding ints 0..3."""
    qt = np.asarray(qt, dtype=np.float32)
    n = qt.shape[0]
    pooled = qt.reshape(n, 4, 2, 4, 2).max(axis=(2, 4))           # F.max_pool2d(.,2)
    pooled = np.clip(np.round(pooled), 0, 3)                      # torch.round == half-to-even
    out = np.stack([square_unity(pooled[i]) for i in range(n)]) if n else pooled
    out = out.repeat(2, axis=1).repeat(2, axis=2)                  # nearest x2
    return out.reshape(n, 1, 8, 8).astype(np.float32)
