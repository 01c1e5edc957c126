"""Work sharding of one sequence over the GPUs of a box (host logic, no CUDA needed).

Blocks are independent end to end and every PartitionMat file is laid out frame by frame (Map2Partition.py:389-412), so a
shard is a contiguous frame range of one QP: the nf x nqp (QP, frame) pairs, taken in QP-major order, are split evenly over
the workers.  Each file is then the concatenation of the rank-ordered segments of its QP.  (Whole frames with all QPs per
worker leave e.g. 30 frames on 8 GPUs at 4 vs 3.75 frames per worker; 120 pairs split 15 each.)
"""


def qp_frame_shards(nf, nqp, world, rank):
    """[(qp_index, frame_lo, frame_hi)] of worker `rank`: at most two QPs unless world < nqp, ranges in QP-major order."""
    total = nf * nqp
    u_lo, u_hi = total * rank // world, total * (rank + 1) // world
    out = []
    for qi in range(nqp):
        a, b = max(u_lo, qi * nf), min(u_hi, (qi + 1) * nf)
        if b > a:
            out.append((qi, a - qi * nf, b - qi * nf))
    return out


def frame_shards(nf, nqp, world, rank):
    """Plain frame ranges with every QP (the reference's DataParallel-style split)."""
    lo, hi = nf * rank // world, nf * (rank + 1) // world
    return [(qi, lo, hi) for qi in range(nqp)] if hi > lo else []


def group_calls(pieces):
    """Merge consecutive pieces with the same frame range: [([qp_index...], frame_lo, frame_hi)] -- one block cut per call."""
    calls = []
    for qi, a, b in pieces:
        if calls and calls[-1][1:] == (a, b):
            calls[-1][0].append(qi)
        else:
            calls.append(([qi], a, b))
    return calls
