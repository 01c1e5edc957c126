"""Drop-in for the reference's ``Map2Partition.py`` entry points: ``map_to_parititon`` (:368-373, sic),
``Map_to_Partition(...).get_partition()`` (:98-365) and ``get_sequence_partition_for_VTM`` (:375-417).
The decode runs on the GPU (one warp per 64x64 block, ``pmp_map2partition``); file content is byte-identical."""
import numpy as np
import torch

from . import ops


def _qt_to_u8(qt_map):
    q = np.asarray(qt_map, dtype=np.float32)
    r = np.clip(q, 0, 4)
    if not np.array_equal(r, np.floor(q)):
        raise ValueError("qt_map must hold integer depths >= 0 (the output of eli_structual_error)")
    return r.astype(np.uint8)


def _decode_batch(qt_map, bt_map, dire_map, chroma_factor, device=None, lamb=None):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    qt = torch.from_numpy(_qt_to_u8(qt_map).reshape(-1, 64)).to(dev)
    bt = torch.from_numpy(np.ascontiguousarray(bt_map, dtype=np.float32).reshape(-1, 3, 16, 16)).to(dev)
    dire = torch.from_numpy(np.ascontiguousarray(dire_map, dtype=np.float32).reshape(-1, 3, 16, 16)).to(dev)
    hor, ver, dout, flags = ops.map2partition(qt, bt, dire, chroma_factor, lamb=lamb)
    return qt, hor, ver, dout, flags


def map_to_parititon(qt_map, bt_map, dire_map, chroma_factor, lamb=None):
    """One block: qt [8,8], bt/dire [3,16,16] -> (hor [16,16] u8, ver [16,16] u8, dire [3,16,16] i8).
    (``lamb``: optional (lamb1..lamb5), an extension -- the reference's wrapper always uses the defaults, :369.)"""
    _, hor, ver, dout, _ = _decode_batch(np.asarray(qt_map)[None], np.asarray(bt_map)[None], np.asarray(dire_map)[None],
                                         chroma_factor, lamb=lamb)
    return hor[0].cpu().numpy(), ver[0].cpu().numpy(), dout[0].cpu().numpy()


class Map_to_Partition:
    """Constructor/``get_partition`` surface of the reference class (returns the [:16,:16] views padded to 17x17)."""

    def __init__(self, qt_map, msbt_map, msdire_map, chroma_factor, lamb1=0.7, lamb2=0.7, lamb3=1.5, lamb4=0.3,
                 lamb5=0.7):
        self._args = (qt_map, msbt_map, msdire_map, chroma_factor)
        self.lamb1, self.lamb2, self.lamb3, self.lamb4, self.lamb5 = lamb1, lamb2, lamb3, lamb4, lamb5   # :118-122

    def get_partition(self):
        hor, ver, dout = map_to_parititon(*self._args, lamb=(self.lamb1, self.lamb2, self.lamb3, self.lamb4, self.lamb5))
        par = np.zeros((2, 17, 17), dtype=np.uint8)
        par[0, :16, :16], par[1, :16, :16] = hor, ver
        return par, dout


def sequence_partition(qt_map, bt_map, dire_map, is_luma, frm_num, frm_width, frm_height, device=None):
    """Device-side body of get_sequence_partition_for_VTM: int8 tensor [frm_num, per-frame values] (file order)."""
    bh, bw = frm_height // 64, frm_width // 64
    n = frm_num * bh * bw
    qt, hor, ver, dout, _ = _decode_batch(np.asarray(qt_map)[:n], np.asarray(bt_map)[:n], np.asarray(dire_map)[:n],
                                          1 if is_luma else 2, device)
    return ops.assemble_frames(hor, ver, qt, dout, frm_num, bh, bw)


def write_partition_text(values, save_path):
    """values: int8 CUDA tensor in file order -> text file (one decimal integer per line, LF)."""
    text = ops.format_text(values)
    with open(save_path, "wb") as f:
        f.write(text.cpu().numpy().tobytes())


def get_sequence_partition_for_VTM(qt_map, bt_map, dire_map, is_luma, save_path, frm_num, frm_width, frm_height):
    vals = sequence_partition(qt_map, bt_map, dire_map, is_luma, frm_num, frm_width, frm_height)
    if save_path is not None:
        write_partition_text(vals, save_path)
