"""lavila_b200.models.utils (checkpoint plumbing) against golden vectors of the unmodified reference
(tests/golden/make_golden_utils.py).  CPU only; exact."""
import os

import pytest
import torch

from lavila_b200.models import utils as U
from tests.golden.make_golden_utils import clip_visual_state

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "utils_small.pt"), weights_only=False)


@pytest.mark.parametrize("idx", range(5))
def test_inflate_positional_embeds_matches_reference(idx):
    c = G["inflate"][idx]
    ckpt = {"visual.temporal_embed": c["input"].clone(), "visual.pos_embed": torch.zeros(1, 5, 12)}
    cur = {"visual.temporal_embed": torch.zeros(1, c["t_model"], 12), "visual.pos_embed": torch.zeros(1, 5, 12)}
    res = U.inflate_positional_embeds(cur, ckpt, num_frames=c["t_model"], load_temporal_fix=c["fix"])
    assert torch.equal(res["visual.temporal_embed"], c["output"])


def test_inflate_refuses_other_patch_counts():
    ckpt = {"visual.pos_embed": torch.zeros(1, 5, 12)}
    cur = {"visual.pos_embed": torch.zeros(1, 10, 12)}
    with pytest.raises(NotImplementedError):
        U.inflate_positional_embeds(cur, ckpt)


def test_remap_keys_matches_reference():
    r = G["remap"]
    sd = clip_visual_state(r["layers"], r["width"], torch.Generator().manual_seed(1))
    rem = U.remap_keys(sd, transformer_layers=r["layers"])
    assert list(rem.keys()) == r["keys"]
    for k, v in rem.items():
        assert tuple(v.shape) == r["shapes"][k]
        assert float(v.double().sum()) == r["checksums"][k]
    with pytest.raises(KeyError):
        U.remap_keys({"unknown.key": torch.zeros(1)})


def test_rsetattr_rgetattr():
    m = torch.nn.Sequential(torch.nn.Linear(2, 2))
    U.rsetattr(m, "0.weight.data", torch.ones(2, 2))
    assert float(U.rgetattr(m, "0.weight").sum()) == 4.0
    assert U.rgetattr(m, "0.missing", None) is None


def test_clip_hf_state_dict_contract():
    """Our CLIP_HF + the HF DistilBERT module exposes exactly the reference's state_dict keys and shapes."""
    import torch.nn as nn
    from transformers import DistilBertConfig, DistilBertModel
    from lavila_b200.models.models import CLIP_HF
    from lavila_b200.models.timesformer import SpaceTimeTransformer, QuickGELU
    vis = SpaceTimeTransformer(img_size=32, patch_size=16, embed_dim=64, depth=1, num_heads=1, num_frames=2, time_init='zeros',
                               ln_pre=True, act_layer=QuickGELU)
    vis.head = nn.Identity()
    vis.pre_logits = nn.Identity()
    m = CLIP_HF(embed_dim=16, vision_width=64, vision_model=vis, text_width=32,
                text_model=DistilBertModel(DistilBertConfig(vocab_size=100, dim=32, n_layers=1, n_heads=2, hidden_dim=64,
                                                            max_position_embeddings=16)),
                text_use_cls_token=True, text_is_regressive=False)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == G["clip_hf_state"]
