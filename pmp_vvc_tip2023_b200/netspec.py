"""Static description of the four Down-Up-CNN nets: parameter names, shapes, FLOPs.

Parameter names and shapes are the reference's ``state_dict`` contract
(/root/reference/Model_QBD.py:59-253; key list probed in SURVEY.md section 8(b)).
This table drives weight loading/packing, seeded-weight generation for tests and
benches (the trained ``*_BD_*.pkl`` are not available offline) and the FLOP
accounting used for the roofline.
"""
from collections import OrderedDict

NETS = ("Luma_Q", "Luma_MSBD", "Chroma_Q", "Chroma_MSBD")
QPS = (22, 27, 32, 37)           # Inference_QBD.py:208


def _rb(prefix, cin, cout, k):
    keys = [(prefix + ".left.0.weight", (cout, cin, k, k)),
            (prefix + ".left.2.weight", (cout, cout, k, k))]
    if cin != cout:
        keys.append((prefix + ".shortcut.0.weight", (cout, cin, 1, 1)))
    return keys


def _q_spec(luma):
    cin, k1, k12 = (1, 9, 5) if luma else (3, 5, 3)
    s = [("conv_q1.weight", (32, cin, k1, k1)), ("conv_q1.bias", (32,))]
    s += _rb("resblock_q1", 32, 64, k12)
    s += _rb("resblock_q2", 64, 64, k12)
    s += _rb("resblock_q3", 64, 32, 3)
    s += _rb("resblock_q4", 128, 32, 3)
    s += _rb("resblock_q5", 32, 32, 3)
    s += _rb("resblock_q6", 32, 8, 3)
    s += [("conv_q2.weight", (1, 8, 3, 3)), ("conv_q2.bias", (1,))]
    return s


def _msbd_spec(luma):
    cin, kb, ks = (2, 9, 5) if luma else (4, 5, 3)
    s = [("conv_b1_1.weight", (16, cin, kb, kb)), ("conv_b1_1.bias", (16,)),
         ("conv_b1_2.weight", (8, cin, ks, kb)), ("conv_b1_2.bias", (8,)),
         ("conv_b1_3.weight", (8, cin, kb, ks)), ("conv_b1_3.bias", (8,))]
    s += _rb("trunk_M1.0", 32, 64, 5)
    for i in range(1, 6):
        s += _rb("trunk_M1.%d" % i, 64, 64, 3)
    for i in range(4):
        s += _rb("trunk_M2.%d" % i, 64, 64, 3)
    for b in ("B1", "B2", "B3"):
        s += _rb("trunk_%s.0" % b, 64, 32, 3)
        s += _rb("trunk_%s.1" % b, 32, 16, 3)
        s += _rb("trunk_%s.2" % b, 16, 8, 3)
    for b in ("B1", "B2", "B3"):
        s += [("conv_%s.weight" % b, (2, 8, 3, 3)), ("conv_%s.bias" % b, (2,))]
    for a in ("Att1", "Att2"):
        s += _rb("trunk_%s.0" % a, 3, 32, 3)
        s += _rb("trunk_%s.1" % a, 32, 64, 3)
    return s


def param_spec(net):
    """Ordered (name, shape) list for ``net`` in the reference's state_dict order."""
    if net == "Luma_Q":
        return _q_spec(True)
    if net == "Chroma_Q":
        return _q_spec(False)
    if net == "Luma_MSBD":
        return _msbd_spec(True)
    if net == "Chroma_MSBD":
        return _msbd_spec(False)
    raise KeyError(net)


def param_count(net):
    n = 0
    for _, shp in param_spec(net):
        c = 1
        for d in shp:
            c *= d
        n += c
    return n


# Algorithmic MACs per 64x64 block (measured with conv
# forward hooks on the reference modules).
MACS_PER_BLOCK = OrderedDict([
    ("Luma_Q", 883.24e6), ("Luma_MSBD", 2612.39e6),
    ("Chroma_Q", 162.05e6), ("Chroma_MSBD", 987.89e6),
])
FLOPS_PER_BLOCK = 2.0 * sum(MACS_PER_BLOCK.values())       # 9.2911 GFLOP
FLOPS_PER_CTU = 4.0 * FLOPS_PER_BLOCK                       # 37.164 GFLOP per 128x128 CTU
