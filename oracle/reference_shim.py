"""Import shims that let the UNMODIFIED reference (`/root/reference/lavila`) import in this
container so that golden vectors can be generated from it (tests/golden/make_golden.py).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path imports this file, and `/root/reference`
does not exist on the GPU box: only the committed fixtures under tests/golden/ travel.

Missing third-party modules: `timm` (DropPath / to_2tuple / trunc_normal_ used at
lavila/models/timesformer.py:31), `ftfy` (lavila/models/tokenizer.py), `decord` (lavila/data).
transformers 5.5 lacks a few legacy symbols that lavila/models/gpt2_gated.py:47-65 and narrator.py:16 import.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LAVILA_REFERENCE_ROOT", "/root/reference")


def _stub(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []
    sys.modules[name] = m
    return m


def install():
    """Install stubs and put the reference on sys.path.  Returns True if the reference is present."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "lavila")):
        return False
    import torch
    import torch.nn as nn

    if "timm" not in sys.modules:
        timm = _stub("timm")
        tm = _stub("timm.models")
        tl = _stub("timm.models.layers")

        class DropPath(nn.Module):
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1 - self.drop_prob
                mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
                return x * mask / keep

        tl.DropPath = DropPath
        tl.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        tl.trunc_normal_ = nn.init.trunc_normal_
        timm.models = tm
        tm.layers = tl
    if "ftfy" not in sys.modules:
        _stub("ftfy").fix_text = lambda s: s
    if "decord" not in sys.modules:
        _stub("decord")

    # transformers 5.x drift (only needed for gpt2_gated / narrator / models.py imports)
    try:
        import transformers
        from transformers import GPT2LMHeadModel, DistilBertModel  # noqa: F401  (settle the lazy module)
        import transformers.modeling_utils as mu
        import transformers.pytorch_utils as pu
        if not hasattr(mu, "SequenceSummary"):
            class SequenceSummary(nn.Module):
                def __init__(self, *a, **k):
                    super().__init__()
            mu.SequenceSummary = SequenceSummary
        for nm in ("find_pruneable_heads_and_indices", "prune_conv1d_layer"):
            if not hasattr(pu, nm):
                def _raise(*a, _n=nm, **k):
                    raise NotImplementedError(_n)
                setattr(pu, nm, _raise)
        if "transformers.utils.model_parallel_utils" not in sys.modules:
            mp = _stub("transformers.utils.model_parallel_utils")
            mp.assert_device_map = lambda *a, **k: None
            mp.get_device_map = lambda *a, **k: {}
        class BeamSearchScorer:  # dummy: beam search is out of scope
            pass
        for mod in {id(transformers): transformers, id(sys.modules["transformers"]): sys.modules["transformers"]}.values():
            if "BeamSearchScorer" not in mod.__dict__:
                mod.__dict__["BeamSearchScorer"] = BeamSearchScorer
        if not hasattr(mu.PreTrainedModel, "get_head_mask"):
            mu.PreTrainedModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
    except Exception as e:  # pragma: no cover
        print("reference_shim: transformers shim incomplete:", e)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    return True
