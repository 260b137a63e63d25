"""Import plumbing for the UNMODIFIED reference installed at baseline/_ref (see baseline/install_ref.py).

BASELINE INFRASTRUCTURE ONLY: used by `bench.py --impl reference`, `bench.py --impl eager` and the in-run
`eager_baseline` / `cpu_baseline` legs.  The product path (lavila_b200/) never imports this file.

What is stubbed, and why: the reference's model files import three symbols from `timm==0.5.4`
(`DropPath`, `to_2tuple`, `trunc_normal_`, lavila/models/timesformer.py:31), `ftfy.fix_text` (tokenizer.py) and `decord`
(lavila/data) -- none is installed here and none does arithmetic on the benchmarked path (drop-path p = 0).  The weight
download inside the factories (`load_openai_clip('ViT-B/16', 'cpu')`, lavila/models/models.py:329) is replaced by a
randomly initialised `lavila.models.openai_model.CLIP` of the same architecture (there is no network); everything after
that line -- `remap_keys`, `load_state_dict`, `CLIP(...)`, the copies of the text tower -- is the reference's own code.
"""
import importlib.machinery
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "lavila", "models", "models.py"))


def _stub(name):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []
    sys.modules[name] = m
    return m


def install():
    """Make `import lavila` resolve to baseline/_ref.  Returns False when the reference is not installed."""
    if not available():
        return False
    import torch.nn as nn

    if "timm" not in sys.modules:
        timm, tm, tl = _stub("timm"), _stub("timm.models"), _stub("timm.models.layers")

        class DropPath(nn.Module):          # stochastic depth; identity at p = 0 / eval (the benchmarked setting)
            def __init__(self, drop_prob=0.0):
                super().__init__()
                self.drop_prob = drop_prob

            def forward(self, x):
                if self.drop_prob == 0.0 or not self.training:
                    return x
                keep = 1 - self.drop_prob
                return x * x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep) / keep

        tl.DropPath, tl.trunc_normal_ = DropPath, nn.init.trunc_normal_
        tl.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
        timm.models, tm.layers = tm, tl
    if "ftfy" not in sys.modules:
        _stub("ftfy").fix_text = lambda s: s
    if "decord" not in sys.modules:
        _stub("decord")
    _transformers_drift(nn)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    return True


def _transformers_drift(nn):
    """lavila/models/gpt2_gated.py:47-65 and narrator.py:16-24 import legacy symbols transformers 5.5 no longer ships;
    lavila/models/models.py imports both files at module level, so the dual-encoder factories need them to import."""
    try:
        import transformers
        from transformers import GPT2LMHeadModel, DistilBertModel  # noqa: F401  (settle the lazy module first)
        import transformers.modeling_utils as mu
        import transformers.pytorch_utils as pu
    except Exception as e:          # pragma: no cover
        print("ref_shim: transformers unavailable:", e, file=sys.stderr)
        return
    if not hasattr(mu, "SequenceSummary"):
        mu.SequenceSummary = type("SequenceSummary", (nn.Module,), {"__init__": lambda self, *a, **k: nn.Module.__init__(self)})
    for nm in ("find_pruneable_heads_and_indices", "prune_conv1d_layer"):
        if not hasattr(pu, nm):
            def _raise(*a, _n=nm, **k):
                raise NotImplementedError(_n)
            setattr(pu, nm, _raise)
    if "transformers.utils.model_parallel_utils" not in sys.modules:
        mp = _stub("transformers.utils.model_parallel_utils")
        mp.assert_device_map = lambda *a, **k: None
        mp.get_device_map = lambda *a, **k: {}
    for mod in {id(transformers): transformers, id(sys.modules["transformers"]): sys.modules["transformers"]}.values():
        if "BeamSearchScorer" not in mod.__dict__:
            mod.__dict__["BeamSearchScorer"] = type("BeamSearchScorer", (), {})
    if not hasattr(mu.PreTrainedModel, "get_head_mask"):
        mu.PreTrainedModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n
