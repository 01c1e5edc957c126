"""Host-side logic of the Model_QBD drop-in modules that needs no GPU: the state_dict contract and the parameter lookup
used to pack weights, including on nn.DataParallel replicas (whose weights are plain tensor attributes)."""
import torch

from pmp_vvc_tip2023_b200 import Model_QBD
from pmp_vvc_tip2023_b200.netspec import param_spec

NETS = (Model_QBD.Luma_Q_Net, Model_QBD.Luma_MSBD_Net, Model_QBD.Chroma_Q_Net, Model_QBD.Chroma_MSBD_Net)


def test_param_list_matches_state_dict_order():
    for cls in NETS:
        net = cls()
        spec = param_spec(net.NET)
        sd = dict(net.named_parameters())
        assert set(sd) == {name for name, _ in spec}            # reference key names, nothing extra
        plist = net._param_list()
        assert len(plist) == len(spec)
        for p, (name, shape) in zip(plist, spec):
            assert p is sd[name] and tuple(p.shape) == tuple(shape)


def _fake_replicate(module):
    """What torch.nn.parallel.replicate leaves behind on the non-source devices: empty _parameters, tensors as attributes."""
    rep = module._replicate_for_data_parallel()
    for key, p in module._parameters.items():
        if p is not None:
            setattr(rep, key, p.detach().clone())
    for key, child in module._modules.items():
        rep._modules[key] = _fake_replicate(child)
    return rep


def test_param_list_on_dataparallel_replica():
    for cls in NETS:
        net = cls()
        rep = _fake_replicate(net)
        assert not dict(rep.named_parameters()) and rep._is_replica
        plist = rep._param_list()
        for p, q in zip(plist, net._param_list()):
            assert isinstance(p, torch.Tensor) and torch.equal(p, q)


def test_weight_cache_fingerprint_is_per_instance():
    """A new net whose parameters land on a dead net's addresses must not reuse that net's packed weights: the cache
    fingerprint carries a never-reused per-instance serial (id() and data_ptr() are both recycled)."""
    from pmp_vvc_tip2023_b200 import Model_QBD
    a = Model_QBD.Chroma_Q_Net()
    fa = a._fingerprint()
    del a
    b = Model_QBD.Chroma_Q_Net()
    assert b._fingerprint()[0] != fa[0]
    g0 = b._fingerprint()
    b.load_state_dict(b.state_dict())
    assert b._fingerprint() != g0            # load_state_dict invalidates the packed weights
