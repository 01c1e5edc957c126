"""N>1 host logic on CPU: frame sharding bounds, per-rank concatenation order == single-process file order, and the
max-over-ranks reduction bench.py uses, with world_size-2 gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import decode_ref
from pmp_vvc_tip2023_b200 import sharding, synth


def shard_bounds(nf, world):
    """Same arithmetic as Inference_QBD.inference_VVC_seqs: contiguous frame ranges per GPU."""
    return [nf * g // world for g in range(world + 1)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nf, bh, bw, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = shard_bounds(nf, world)
    lo, hi = b[rank], b[rank + 1]
    qt, bt, dire = synth.structured_maps(nf * bh * bw, seed=5, sigma=0.1)
    sl = slice(lo * bh * bw, hi * bh * bw)
    hor, ver, qtm, dm = decode_ref.sequence_partition(qt[sl], bt[sl], dire[sl], True, hi - lo, bw * 64, bh * 64)
    with open(os.path.join(out_dir, "seg%d.txt" % rank), "wb") as f:
        f.write(decode_ref.partition_text(hor, ver, qtm, dm))
    # bench.py timing rule: max over ranks
    t = torch.tensor([10.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == 10.0 + world - 1
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for nf in (1, 2, 7, 10, 30):
        for world in (1, 2, 4, 8):
            b = shard_bounds(nf, world)
            assert b[0] == 0 and b[-1] == nf and all(b[i] <= b[i + 1] for i in range(world))


def test_two_rank_segments_concatenate_to_single_process_file(tmp_path):
    nf, bh, bw = 3, 1, 2
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nf, bh, bw, str(tmp_path)), nprocs=world, join=True)
    qt, bt, dire = synth.structured_maps(nf * bh * bw, seed=5, sigma=0.1)
    hor, ver, qtm, dm = decode_ref.sequence_partition(qt, bt, dire, True, nf, bw * 64, bh * 64)
    want = decode_ref.partition_text(hor, ver, qtm, dm)
    got = b"".join(open(os.path.join(str(tmp_path), "seg%d.txt" % r), "rb").read() for r in range(world))
    assert got == want


def test_qp_frame_shards_cover_each_file_in_rank_order():
    """(QP, frame) pair sharding: per QP the rank-ordered segments tile [0, nf) exactly, work differs by at most one pair."""
    for nf in (1, 2, 7, 10, 30, 33):
        for world in (1, 2, 3, 4, 8):
            loads = []
            cover = {qi: [] for qi in range(4)}
            for r in range(world):
                pieces = sharding.qp_frame_shards(nf, 4, world, r)
                loads.append(sum(b - a for _, a, b in pieces))
                for qi, a, b in pieces:
                    cover[qi].append((a, b))
                assert sum(len(qs) * (b - a) for qs, a, b in sharding.group_calls(pieces)) == loads[-1]
            assert sum(loads) == 4 * nf and max(loads) - min(loads) <= 1
            for qi, segs in cover.items():
                assert segs[0][0] == 0 and segs[-1][1] == nf
                assert all(segs[i][1] == segs[i + 1][0] for i in range(len(segs) - 1))
    # the plain frame split shares one call between all QPs
    assert sharding.group_calls(sharding.frame_shards(30, 4, 8, 1)) == [([0, 1, 2, 3], 3, 7)]
