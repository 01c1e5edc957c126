"""Plain-PyTorch CPU fp32 restatement of the four Down-Up-CNN forwards (TEST ORACLE ONLY).

Follows /root/reference/Model_QBD.py:
  ResidualBlock.forward   :40-44
  Luma_Q_Net.forward      :78-98
  Luma_MSBD_Net.forward   :127-155
  Chroma_Q_Net.forward    :176-196
  Chroma_MSBD_Net.forward :225-253
and the batch regrouping of Metrics.inference_pre_QBD (:387-419).

Functional form over a ``state_dict`` that uses the reference parameter names
(``module.`` prefix already stripped).  Pinned against the reference modules
themselves by tests/golden/nets_*.npz (Q nets with the reference's trained
weights; MSBD nets with seeded random weights because the trained *_BD_*.pkl are
absent from the mount -- "parity with trained MTT weights unpinned").
"""
import numpy as np
import torch
import torch.nn.functional as F


def _resblock(sd, prefix, x, k):
    """Model_QBD.py:23-44 -- relu(conv2(relu(conv1 x)) + shortcut(x)); no bias anywhere."""
    p = k // 2
    out = F.conv2d(x, sd[prefix + ".left.0.weight"], padding=p)
    out = F.relu(out)
    out = F.conv2d(out, sd[prefix + ".left.2.weight"], padding=p)
    key = prefix + ".shortcut.0.weight"
    out = out + (F.conv2d(x, sd[key]) if key in sd else x)
    return F.relu(out)


def _trunk(sd, prefix, x, ks):
    for i, k in enumerate(ks):
        x = _resblock(sd, "%s.%d" % (prefix, i), x, k)
    return x


def q_net_forward(sd, x, luma):
    """Luma_Q_Net (:78-98) / Chroma_Q_Net (:176-196).  x: [B,1,68,68] or [B,3,34,34]."""
    ov = 4 if luma else 2
    k12 = 5 if luma else 3
    x1 = F.pad(x, (0, ov, 0, ov))                                           # padding_rb
    x2 = F.relu(F.conv2d(x1, sd["conv_q1.weight"], sd["conv_q1.bias"]))
    x3 = _resblock(sd, "resblock_q1", x2, k12)
    if luma:
        x3 = F.max_pool2d(x3, 2)
    x4 = F.max_pool2d(_resblock(sd, "resblock_q2", x3, k12), 2)
    x5 = _resblock(sd, "resblock_q3", x4, 3)
    pyr = [x5] + [F.interpolate(F.max_pool2d(x5, s), scale_factor=s) for s in (2, 4, 8)]
    x6 = torch.cat(pyr, 1)
    x7 = _resblock(sd, "resblock_q4", x6, 3)
    x8 = F.max_pool2d(_resblock(sd, "resblock_q5", x7, 3), 2)
    x9 = _resblock(sd, "resblock_q6", x8, 3)
    return F.conv2d(x9, sd["conv_q2.weight"], sd["conv_q2.bias"], padding=1)


def msbd_net_forward(sd, x, qt, luma):
    """Luma_MSBD_Net (:127-155) / Chroma_MSBD_Net (:225-253).  qt: raw float [B,1,8,8]."""
    ov = 4 if luma else 2
    up = 8 if luma else 4
    x1_1 = F.pad(F.interpolate(qt, scale_factor=up), (ov, 0, ov, 0))       # padding_lu
    x2 = torch.cat([x, x1_1], 1)
    x3_1 = F.relu(F.conv2d(F.pad(x2, (0, ov, 0, ov)), sd["conv_b1_1.weight"], sd["conv_b1_1.bias"]))
    x3_2 = F.relu(F.conv2d(F.pad(x2, (0, ov, 0, 0)), sd["conv_b1_2.weight"], sd["conv_b1_2.bias"]))
    x3_3 = F.relu(F.conv2d(F.pad(x2, (0, 0, 0, ov)), sd["conv_b1_3.weight"], sd["conv_b1_3.bias"]))
    x3 = torch.cat([x3_1, x3_2, x3_3], 1)
    x4 = _trunk(sd, "trunk_M1", x3, (5, 3, 3, 3, 3, 3))
    if luma:
        x4 = F.max_pool2d(x4, 2)
    x5 = F.max_pool2d(_trunk(sd, "trunk_M2", x4, (3, 3, 3, 3)), 2)
    x6 = _trunk(sd, "trunk_B1", x5, (3, 3, 3))
    out0 = F.conv2d(x6, sd["conv_B1.weight"], sd["conv_B1.bias"], padding=1)
    out0q = torch.cat([F.interpolate(qt, scale_factor=2), out0], 1)
    att0 = _trunk(sd, "trunk_Att1", out0q, (3, 3))
    xb2 = _trunk(sd, "trunk_B2", x5 * att0, (3, 3, 3))
    out1 = F.conv2d(xb2, sd["conv_B2.weight"], sd["conv_B2.bias"], padding=1)
    out1 = torch.cat([out1[:, 0:1] + out0[:, 0:1], out1[:, 1:2]], 1)       # :146 (in place there)
    out1q = torch.cat([F.interpolate(qt, scale_factor=4), F.interpolate(out1, scale_factor=2)], 1)
    att1 = _trunk(sd, "trunk_Att2", out1q, (3, 3))
    xb4 = F.max_pool2d(_trunk(sd, "trunk_B3", x4 * att1, (3, 3, 3)), 2)
    out2 = F.conv2d(xb4, sd["conv_B3.weight"], sd["conv_B3.bias"], padding=1)
    out2 = torch.cat([out2[:, 0:1] + out1[:, 0:1], out2[:, 1:2]], 1)       # :153
    return out0, out1, out2


@torch.no_grad()
def predict_maps(sd_q, sd_bd, blocks, luma, batch=200):
    """Metrics.inference_pre_QBD (:387-419): Q -> MSBD(raw qt) -> regroup.

    blocks: float32 tensor [N,1,68,68] (luma) or [N,3,34,34] (chroma).
    Returns (qt[N,1,8,8], bt[N,3,16,16], dire[N,3,16,16]) float32 CPU tensors."""
    qs, bs, ds = [], [], []
    for i in range(0, blocks.shape[0], batch):
        x = blocks[i:i + batch]
        qt = q_net_forward(sd_q, x, luma)
        o0, o1, o2 = msbd_net_forward(sd_bd, x, qt, luma)
        qs.append(qt)
        bs.append(torch.cat([o0[:, 0:1], o1[:, 0:1], o2[:, 0:1]], 1))
        ds.append(torch.cat([o0[:, 1:2], o1[:, 1:2], o2[:, 1:2]], 1))
    return torch.cat(qs), torch.cat(bs), torch.cat(ds)


def chroma_net_input(block_y, block_u, block_v):
    """Inference_QBD.py:194-200 -- cat(max_pool2d(luma 68x68, 2), U, V) -> [N,3,34,34] float32."""
    y = torch.as_tensor(np.asarray(block_y), dtype=torch.float32).unsqueeze(1)
    u = torch.as_tensor(np.asarray(block_u), dtype=torch.float32).unsqueeze(1)
    v = torch.as_tensor(np.asarray(block_v), dtype=torch.float32).unsqueeze(1)
    return torch.cat([F.max_pool2d(y, 2), u, v], 1)


def cut_blocks(y, u, v, is10bit):
    """Inference_QBD.output_block_yuv (:104-149) on in-memory planes.

    y [F,H,W], u/v [F,H/2,W/2] (uint16 if is10bit else uint8).  Returns
    block_y [N,68,68], block_u/v [N,34,34] uint8, frame-major raster order."""
    outs = []
    for idx, comp in enumerate((y, u, v)):
        comp = np.asarray(comp)
        if is10bit:
            comp = np.round(comp / 4).clip(0, 255).astype(np.uint8)        # half-to-even (:106-109)
        ov, bs = (4, 64) if idx == 0 else (2, 32)
        pad = np.zeros((comp.shape[0], comp.shape[1] + ov, comp.shape[2] + ov), np.uint8)
        pad[:, ov:, ov:] = comp
        nbh, nbw = y.shape[1] // 64, y.shape[2] // 64
        blk = [pad[f, i * bs:(i + 1) * bs + ov, j * bs:(j + 1) * bs + ov]
               for f in range(comp.shape[0]) for i in range(nbh) for j in range(nbw)]
        outs.append(np.array(blk))
    return outs
