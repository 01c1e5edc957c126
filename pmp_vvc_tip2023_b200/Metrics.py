"""Drop-in for the hot-path functions of the reference's ``Metrics.py``:
``inference_pre_QBD`` (:387-419), ``check_square_unity`` (:612-628), ``eli_structual_error`` (:630-637) and
``seq_post_process`` (:764-774).  Same names, argument meaning and return types; the compute is libpmp_b200."""
import numpy as np
import torch

from . import ops
from .Map2Partition import get_sequence_partition_for_VTM


@torch.no_grad()
def inference_pre_QBD(infe_loader_QB, Net_Q, Net_BD):
    """Batch loop: Q net -> MSBD net (fed the raw, un-rounded qt) -> regroup.  Returns CPU float tensors
    (qt [N,1,8,8], bt [N,3,16,16], dire [N,3,16,16]) like the reference (one concat at the end instead of one
    growing concat per batch)."""
    qts, bts, dires = [], [], []
    for data in infe_loader_QB:
        input_batch = data[0].cuda(non_blocking=True)
        qt = Net_Q(input_batch)
        o0, o1, o2 = Net_BD(input_batch, qt)
        qts.append(qt.cpu())
        bts.append(torch.cat([o0[:, 0:1], o1[:, 0:1], o2[:, 0:1]], 1).cpu())
        dires.append(torch.cat([o0[:, 1:2], o1[:, 1:2], o2[:, 1:2]], 1).cpu())
    if not qts:
        return torch.zeros((0, 1, 8, 8)), torch.zeros((0, 3, 16, 16)), torch.zeros((0, 3, 16, 16))
    return torch.cat(qts, 0), torch.cat(bts, 0), torch.cat(dires, 0)


def eli_structual_error(out_batch):
    """[N,1,8,8] float tensor (CUDA) -> same shape holding the repaired integer QT depths 0..3 (sic spelling)."""
    if not out_batch.is_cuda:
        out_batch = out_batch.cuda()          # the reference hard-codes .cuda() here (Metrics.py:615)
    of, _ = ops.qt_postprocess(out_batch, want_f32=True, want_u8=False)
    return of


def check_square_unity(mat):
    """4x4 tensor of rounded/clamped depths -> repaired 4x4 (Metrics.py:612-628), via the batch kernel."""
    m = torch.as_tensor(mat, dtype=torch.float32).cuda().reshape(1, 1, 4, 4)
    up = m.repeat_interleave(2, 2).repeat_interleave(2, 3)      # max_pool2d(up2(m)) == m, round/clamp are no-ops
    return eli_structual_error(up)[0, 0, ::2, ::2].contiguous()


def seq_post_process(input_qt_batch, input_bt_batch, input_dire_batch, comp, sub_numfrm, width, height, save_path):
    is_luma = (comp == "Luma")
    qt = eli_structual_error(torch.as_tensor(input_qt_batch)).cpu().numpy().squeeze(axis=1)
    get_sequence_partition_for_VTM(qt_map=qt, bt_map=np.asarray(input_bt_batch), dire_map=np.asarray(input_dire_batch),
                                   is_luma=is_luma, save_path=save_path, frm_num=sub_numfrm, frm_width=width,
                                   frm_height=height)
