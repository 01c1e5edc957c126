"""NumPy restatement of the reference map-to-partition decode (TEST ORACLE ONLY).

Follows /root/reference/Map2Partition.py:
  th_round                      :30-35
  Search (cross product)        :53-87
  Map_to_Partition.__init__     :100-122
  split_cur_map                 :124-138
  can_split_mode_list           :140-201
  get_candidate_map_tree        :203-266
  set_bt_partition_vector       :287-346
  set_partition_vector          :348-362
  map_to_parititon              :368-373
  get_sequence_partition_for_VTM:375-417

The algorithm is restated, not copied: the reference materialises a tree of
Map_Node objects and a Search tree for the cross product; here the same leaves
are visited in the same order by a recursive generator, and the leaf error is
evaluated with exactly the same NumPy expressions (float32 ``np.sum`` over the
region slice per level, then ``0.8 * (...)`` under NEP-50 scalar rules), so the
first-minimum choice is bit-identical with the reference under NumPy >= 2.

Pinned against the reference itself by tests/golden/decode_*.npz
(see oracle/gen_golden.py).
"""
import itertools

import numpy as np

LAMB1, LAMB2, LAMB3, LAMB4, LAMB5 = 0.7, 0.7, 1.5, 0.3, 0.7   # Map2Partition.py:100
DIRE_WEIGHT = 0.8                                              # Map2Partition.py:310


def threshold_direction(d, thd=0.5):
    """Map2Partition.py:30-35 -- +1 if >= thd, -1 if <= -thd, else 0 (same dtype)."""
    out = np.zeros_like(d)
    out[d >= thd] = 1
    out[d <= -thd] = -1
    return out


def split_rects(x, y, h, w, mode):
    """Map2Partition.py:124-138 -- child rectangles (x=row, y=col, in 4x4-luma cells)."""
    if mode == 0:
        return [(x, y, h, w)]
    if mode == 1:   # BT horizontal
        return [(x, y, h // 2, w), (x + h // 2, y, h // 2, w)]
    if mode == 2:   # BT vertical
        return [(x, y, h, w // 2), (x, y + w // 2, h, w // 2)]
    if mode == 3:   # TT horizontal
        return [(x, y, h // 4, w), (x + h // 4, y, h // 2, w), (x + (h * 3) // 4, y, h // 4, w)]
    if mode == 4:   # TT vertical
        return [(x, y, h, w // 4), (x, y + w // 4, h, w // 2), (x, y + (w * 3) // 4, h, w // 4)]
    raise ValueError(mode)


def child_depth_increment(mode, idx):
    """+1 for every BT/TT child, +2 for the outer TT children (Map2Partition.py:183-186,259-262)."""
    if mode == 0:
        return 0
    return 2 if (mode >= 3 and idx != 1) else 1


class DecodeRef:
    def __init__(self, qt_map, msbt_map, msdire_map, chroma_factor, lamb=(LAMB1, LAMB2, LAMB3, LAMB4, LAMB5)):
        self.lamb1, self.lamb2, self.lamb3, self.lamb4, self.lamb5 = lamb       # :118-122
        self.qt = qt_map
        self.ori_bt = msbt_map
        self.ori_dire = msdire_map
        self.rbt = np.round(msbt_map)                    # :104 half-to-even, no clamp
        self.rdire = threshold_direction(msdire_map)     # :105
        self.cf = chroma_factor
        self.par_vec = np.zeros((2, 17, 17), dtype=np.uint8)   # :112
        self.out_dire = np.zeros((3, 16, 16), dtype=np.int8)   # :114

    # --- :140-201 ---------------------------------------------------------
    def candidate_modes(self, x, y, h, w, cur_bt, d):
        cmp2 = self.rbt[2, x:x + h, y:y + w] - cur_bt[x:x + h, y:y + w]
        if np.count_nonzero(cmp2 == 0) >= self.lamb1 * h * w:
            return [0]
        reg = self.rdire[d, x:x + h, y:y + w]
        n_hor = np.count_nonzero(reg == 1)
        n_ver = np.count_nonzero(reg == -1)
        direction = 0
        if (n_ver + n_hor) >= self.lamb2 * h * w:
            if n_hor >= self.lamb3 * n_ver:
                direction = 1
            elif n_ver >= self.lamb3 * n_hor:
                direction = 2
        cf = self.cf
        kept = [0]
        for mode in (1, 2, 3, 4):
            length = h if mode in (1, 3) else w
            unit = (2 if mode <= 2 else 4) * cf
            if length // unit == 0 or length % unit != 0:
                continue
            if mode in (1, 3) and direction == 2:
                continue
            if mode in (2, 4) and direction == 1:
                continue
            ok = True
            for idx, (sx, sy, sh, sw) in enumerate(split_rects(x, y, h, w, mode)):
                tmp = cur_bt[sx:sx + sh, sy:sy + sw] + child_depth_increment(mode, idx)
                cmpd = self.rbt[d, sx:sx + sh, sy:sy + sw] - tmp
                n_minus = np.count_nonzero(cmpd < 0)
                n_zero = np.count_nonzero(cmpd == 0)
                n = sh * sw
                if not (n_minus < n * self.lamb4 and n_zero > n * self.lamb5):
                    ok = False          # reference keeps counting; result is the same
            if ok:
                kept.append(mode)
        return kept

    # --- :203-266, leaves in the reference's DFS order ----------------------
    def _leaves(self, depth, cus, bt_map, chain):
        """Yield (chain_of_(bt_map,dire_map) per level, cus) for every depth-3 leaf."""
        if depth == 3:
            yield chain, cus
            return
        cand = [self.candidate_modes(cx, cy, ch, cw, bt_map, depth) for (cx, cy, ch, cw) in cus]
        for modes in itertools.product(*cand):       # first CU most significant == Search DFS (:53-87)
            child_bt = bt_map.copy()
            child_dire = np.zeros((16, 16), dtype=np.int8)    # zero-initialised per node (:240)
            child_cus = []
            for (cx, cy, ch, cw), m in zip(cus, modes):
                rects = split_rects(cx, cy, ch, cw, m)
                child_cus += rects
                if m == 0:
                    continue
                child_dire[cx:cx + ch, cy:cy + cw] = 1 if m in (1, 3) else -1
                for idx, (sx, sy, sh, sw) in enumerate(rects):
                    child_bt[sx:sx + sh, sy:sy + sw] += child_depth_increment(m, idx)
            yield from self._leaves(depth + 1, child_cus, child_bt, chain + [(child_bt, child_dire)])

    # --- :287-346 -----------------------------------------------------------
    def decode_mtt_region(self, x, y, h, w):
        zero = np.zeros((16, 16), dtype=np.int8)
        best = None
        best_err = None
        sl = (slice(x, x + h), slice(y, y + w))
        for chain, cus in self._leaves(0, [(x, y, h, w)], zero, []):
            (b0, d0), (b1, d1), (b2, d2) = chain
            err = np.sum(np.abs(b0[sl] - self.ori_bt[0][sl])) + \
                np.sum(np.abs(b1[sl] - self.ori_bt[1][sl])) + \
                np.sum(np.abs(b2[sl] - self.ori_bt[2][sl])) + \
                DIRE_WEIGHT * (np.sum(np.abs(d0[sl] - self.ori_dire[0][sl])) +
                               np.sum(np.abs(d1[sl] - self.ori_dire[1][sl])) +
                               np.sum(np.abs(d2[sl] - self.ori_dire[2][sl])))
            if best is None or err < best_err:          # first minimum (:315)
                best, best_err = (chain, cus), err
        chain, cus = best
        for lvl in range(3):
            self.out_dire[lvl][sl] = chain[lvl][1][sl]
        for (cx, cy, ch, cw) in cus:                    # :339-346
            self.par_vec[0, cx, cy:cy + cw] = 1
            self.par_vec[0, cx + ch, cy:cy + cw] = 1
            self.par_vec[1, cx:cx + ch, cy] = 1
            self.par_vec[1, cx:cx + ch, cy + cw] = 1

    # --- :348-362 -----------------------------------------------------------
    def decode_qt(self, depth, qx, qy):
        cur = self.qt[qx, qy]
        size = 8 >> depth
        if cur == depth:
            self.decode_mtt_region(2 * qx, 2 * qy, 2 * size, 2 * size)
        elif cur > depth:
            self.par_vec[0, 2 * qx + size, 2 * qy:2 * qy + 2 * size] = 1
            self.par_vec[1, 2 * qx:2 * qx + 2 * size, 2 * qy + size] = 1
            for io in range(2):
                for jo in range(2):
                    self.decode_qt(depth + 1, qx + io * size // 2, qy + jo * size // 2)
        # cur < depth: nothing (region left unset)

    def run(self):
        self.decode_qt(0, 0, 0)
        return self.par_vec[0][:16, :16], self.par_vec[1][:16, :16], self.out_dire


def map_to_partition(qt_map, bt_map, dire_map, chroma_factor, lamb=(LAMB1, LAMB2, LAMB3, LAMB4, LAMB5)):
    """Restates Map2Partition.map_to_parititon (:368-373); ``lamb`` = the Map_to_Partition constructor thresholds (:100).

    qt_map [8,8] (integers 0..3 in any dtype), bt_map/dire_map [3,16,16] float32.
    Returns (hor[16,16] u8, ver[16,16] u8, dire[3,16,16] i8)."""
    return DecodeRef(qt_map, bt_map, dire_map, chroma_factor, lamb).run()


def sequence_partition(qt_map, bt_map, dire_map, is_luma, frm_num, frm_width, frm_height):
    """Frame assembly of get_sequence_partition_for_VTM (:375-399), no file I/O.

    Returns (hor[F,R,C] u8, ver[F,R,C] u8, qt[F,R/2,C/2] u8, dire[F,3,R,C] i8)."""
    cf = 1 if is_luma else 2
    bh, bw = frm_height // 64, frm_width // 64
    hor = np.zeros((frm_num, bh * 16, bw * 16), np.uint8)
    ver = np.zeros_like(hor)
    qt = np.zeros((frm_num, bh * 8, bw * 8), np.uint8)
    dire = np.zeros((frm_num, 3, bh * 16, bw * 16), np.int8)
    for f in range(frm_num):
        for bx in range(bh):
            for by in range(bw):
                bid = (f * bh + bx) * bw + by
                h, v, d = map_to_partition(qt_map[bid], bt_map[bid], dire_map[bid], cf)
                hor[f, bx * 16:(bx + 1) * 16, by * 16:(by + 1) * 16] = h
                ver[f, bx * 16:(bx + 1) * 16, by * 16:(by + 1) * 16] = v
                qt[f, bx * 8:(bx + 1) * 8, by * 8:(by + 1) * 8] = qt_map[bid]
                dire[f, :, bx * 16:(bx + 1) * 16, by * 16:(by + 1) * 16] = d
    return hor, ver, qt, dire


def partition_text(hor, ver, qt, dire):
    """File body of get_sequence_partition_for_VTM (:400-412): one decimal integer per line,
    per frame hor | ver | qt | dire, LF endings."""
    parts = []
    for f in range(hor.shape[0]):
        for vec in (hor[f].reshape(-1).astype(np.uint8), ver[f].reshape(-1).astype(np.uint8),
                    qt[f].reshape(-1).astype(np.uint8), dire[f].reshape(-1).astype(np.int8)):
            parts.append("".join("%d\n" % int(v) for v in vec))
    return "".join(parts).encode("ascii")
