"""Weight files: the reference's ``.pkl`` contract (Inference_QBD.py:28-46).

* files are legacy ``torch.save`` pickles of an OrderedDict (optionally wrapped in
  ``{"state_dict": ...}``), keys prefixed ``module.`` (saved from DataParallel), storages
  tagged CUDA -> always load with ``map_location='cpu'`` (the reference's own loader
  raises on a CPU-only host);
* ``load_pretrain_model`` keeps only keys present in the model with a matching shape and
  silently skips the rest.
"""
import torch


def remove_prefix(state_dict, prefix):
    """Inference_QBD.py:28-30."""
    return {(k.split(prefix, 1)[-1] if k.startswith(prefix) else k): v for k, v in state_dict.items()}


def load_reference_pkl(path):
    """Read a reference ``.pkl`` into ``{name: cpu float tensor}`` with ``module.`` stripped."""
    src = torch.load(path, map_location="cpu", weights_only=False)
    if "state_dict" in src.keys():
        src = src["state_dict"]
    return remove_prefix(src, "module.")


def load_pretrain_model(current_model, pretrain_model):
    """Drop-in for Inference_QBD.load_pretrain_model (:33-46): shape-matched partial load."""
    source = load_reference_pkl(pretrain_model) if isinstance(pretrain_model, str) else \
        remove_prefix(pretrain_model, "module.")
    dest = current_model.state_dict()
    trained = {k: v for k, v in source.items() if k in dest and tuple(v.shape) == tuple(dest[k].shape)}
    dest.update(trained)
    current_model.load_state_dict(dest)
    return current_model
