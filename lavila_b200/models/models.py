"""CLIP dual encoder + factories -- mirror of lavila/models/models.py:75-173, 293-313, 316-491 (hot-path subset).

The factories build the same architectures as the reference's CLIP_OPENAI_TIMESFORMER_* functions.  The reference
downloads OpenAI CLIP weights there (models.py:329); without network the model is randomly initialised with the
reference's initialisers, or loaded from a LaViLa/ours checkpoint via `checkpoint=` (state_dict names are identical).
"""
import numpy as np
import torch
import torch.nn as nn

from .. import engine as E
from . import loss as loss_mod
from . import loss  # noqa: F401  main_pretrain.py:189 reaches SSLCLIPLoss as `models.loss.SSLCLIPLoss`
from .openai_model import QuickGELU, Transformer
from .timesformer import SpaceTimeTransformer


class CLIP(nn.Module):
    """models.py:75-173."""

    def __init__(self, embed_dim: int, vision_width: int, vision_model: nn.Module, context_length: int,
                 vocab_size: int, transformer_width: int, transformer_heads: int, transformer_layers: int,
                 tempearture_init=0.07, **kwargs):
        super().__init__()
        self.context_length = context_length
        self.vision_width = vision_width
        self.visual = vision_model
        self.transformer = Transformer(width=transformer_width, layers=transformer_layers, heads=transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = vocab_size
        self.token_embedding = nn.Embedding(vocab_size, transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, transformer_width))
        self.ln_final = nn.LayerNorm(transformer_width)
        self.image_projection = nn.Parameter(torch.empty(vision_width, embed_dim))
        self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim))
        print("=> initialize initial temperature with {}".format(tempearture_init))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / tempearture_init))
        self.initialize_parameters()

    def half(self):
        """`--use-half` (main_infer_narrator.py:155-156, eval_zeroshot.py:142): a no-op here.  The kernels own the precision
        (bf16 operands, fp32 accumulation, fp32 parameters / residual stream / softmax -- at least the accuracy of the reference's
        fp16 module), so the parameters stay fp32 and half-precision inputs are widened on entry."""
        return self

    def bfloat16(self):
        return self

    def initialize_parameters(self):
        """models.py:115-129."""
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.image_projection, std=self.vision_width ** -0.5)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        """models.py:131-137 (kept for API parity; the causal structure is built into the attention kernel)."""
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    def encode_image(self, image, use_checkpoint=False, apply_project=True):
        x = self.visual(image, use_checkpoint=use_checkpoint)
        if isinstance(x, list):
            assert len(x) == 1
            x = x[0]
        if not apply_project:
            return x
        return E.ProjectFn.apply(x, self.image_projection)

    def encode_text(self, text, use_checkpoint=False):
        x = E.TextEmbedFn.apply(text, self.token_embedding.weight, self.positional_embedding)   # [B, L, W]
        x = self.transformer.forward_bld(x)
        x = E.GatherEotFn.apply(x, text)                                                        # EOT rows [B, W]
        x = E.LayerNormFn.apply(x, self.ln_final.weight, self.ln_final.bias, float(self.ln_final.eps))
        return E.ProjectFn.apply(x, self.text_projection)

    def forward(self, image, text, use_checkpoint=False, norm_embed=False):
        image_embed = self.encode_image(image, use_checkpoint=use_checkpoint)
        text_embed = self.encode_text(text, use_checkpoint=use_checkpoint)
        if norm_embed:
            image_embed = E.L2NormalizeFn.apply(image_embed)
            text_embed = E.L2NormalizeFn.apply(text_embed)
        return {'image_embed': image_embed, 'text_embed': text_embed, 'logit_scale': self.logit_scale.exp()}


class CLIP_HF(nn.Module):
    """models.py:176-290: dual encoder whose text tower is a Hugging Face module (DistilBERT in the TSF-L@HR recipe,
    docs/PRETRAIN.md:24-36).  The video tower, both projections, the normalisation and the loss run on the B200 kernels;
    the HF text module (0.8 % of the step's FLOPs) is called as it is -- library code, like the reference does -- and its
    output joins the kernel path through ordinary autograd."""

    def __init__(self, embed_dim: int, vision_width: int, vision_model: nn.Module, text_width: int, text_model: nn.Module,
                 text_use_cls_token: bool, text_is_regressive: bool, tempearture_init=0.07, **kwargs):
        super().__init__()
        self.vision_width = vision_width
        self.visual = vision_model
        self.text_width = text_width
        self.textual = text_model
        self.text_use_cls_token = text_use_cls_token
        self.text_is_regressive = text_is_regressive
        self.projection = kwargs.get('projection', 'default')
        if self.projection == 'default':
            self.image_projection = nn.Parameter(torch.empty(vision_width, embed_dim))
            self.text_projection = nn.Parameter(torch.empty(text_width, embed_dim))
        elif self.projection == 'frozen_in_time':
            self.image_projection = nn.Sequential(nn.Linear(vision_width, embed_dim))
            self.text_projection = nn.Sequential(nn.ReLU(), nn.Linear(text_width, embed_dim))
        else:
            raise NotImplementedError(self.projection)
        print("=> initialize initial temperature with {}".format(tempearture_init))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / tempearture_init))
        self.initialize_parameters()

    def half(self):
        """`--use-half` (main_infer_narrator.py:155-156, eval_zeroshot.py:142): a no-op here.  The kernels own the precision
        (bf16 operands, fp32 accumulation, fp32 parameters / residual stream / softmax -- at least the accuracy of the reference's
        fp16 module), so the parameters stay fp32 and half-precision inputs are widened on entry."""
        return self

    def bfloat16(self):
        return self

    def initialize_parameters(self):
        """models.py:219-225."""
        if self.projection == 'default':
            nn.init.normal_(self.image_projection, std=self.vision_width ** -0.5)
            nn.init.normal_(self.text_projection, std=self.text_width ** -0.5)
        else:
            nn.init.normal_(self.image_projection[0].weight, std=self.vision_width ** -0.5)
            nn.init.normal_(self.text_projection[1].weight, std=self.text_width ** -0.5)

    @staticmethod
    def _linear(x, lin):
        """The 'frozen_in_time' projection heads (EgoVLP, models.py:660-688): [B, K] x [K, N] on B rows -- plain library call."""
        return torch.nn.functional.linear(x, lin.weight, lin.bias)

    def encode_image(self, image, use_checkpoint=False, apply_project=True):
        x = self.visual(image, use_checkpoint=use_checkpoint)
        if isinstance(x, list):
            assert len(x) == 1
            x = x[0]
        if not apply_project:
            return x
        if self.projection == 'default':
            return E.ProjectFn.apply(x, self.image_projection)
        return self._linear(x, self.image_projection[0])

    def encode_text(self, text, attention_mask=None, use_checkpoint=False):
        """models.py:249-280 (the gradient-checkpointing toggles are forwarded when the HF module has them)."""
        toggle = getattr(self.textual, 'gradient_checkpointing_enable' if use_checkpoint else 'gradient_checkpointing_disable', None)
        if toggle is not None and not (use_checkpoint and type(self.textual).__name__ == 'DistilBertModel'):
            try:
                toggle()
            except Exception:
                pass
        x = self.textual(text, attention_mask=attention_mask)
        if self.text_is_regressive:
            x = x.last_hidden_state
            x = x[torch.arange(x.shape[0], device=x.device), text.argmax(dim=-1)]
        elif self.text_use_cls_token:
            x = x.last_hidden_state[:, 0, :]
        else:
            x = x.pooler_output
        x = x.float().contiguous()
        if self.projection == 'default':
            return E.ProjectFn.apply(x, self.text_projection)
        return self._linear(torch.relu(x), self.text_projection[1])

    def forward(self, image, text, mask=None, use_checkpoint=False, norm_embed=False):
        image_embed = self.encode_image(image, use_checkpoint=use_checkpoint)
        text_embed = self.encode_text(text, attention_mask=mask, use_checkpoint=use_checkpoint)
        if norm_embed:
            image_embed = E.L2NormalizeFn.apply(image_embed)
            text_embed = E.L2NormalizeFn.apply(text_embed)
        return {'image_embed': image_embed, 'text_embed': text_embed, 'logit_scale': self.logit_scale.exp()}


def get_loss(model, args, tokenizer=None):
    """models.py:293-304 (CLIP branch)."""
    if model.startswith('CLIP'):
        return loss_mod.CLIPLoss(use_vissl=args.contrastive_use_vissl, cache_labels=True, rank=args.rank,
                                 world_size=args.world_size)
    raise NotImplementedError("only the CLIP* dual-encoder losses are on the B200 hot path")


def get_metric_names(model):
    """models.py:307-313 (the SSLCLIPLoss run adds clip_acc_gt / clip_acc_pseudo itself, main_pretrain.py:235-236)."""
    if model.startswith('CLIP'):
        return ['loss', 'clip_loss', 'clip_acc']
    raise NotImplementedError


_BENIGN_MISSING = ("attn_mask", "attn.bias", "attn.masked_bias", "position_ids")   # buffers rebuilt by the constructors


def _load_checkpoint(model, checkpoint, strict=True):
    """Reference-format checkpoints are `{'epoch', 'state_dict', 'optimizer', 'scaler', 'best_acc1', 'args'}`
    (main_pretrain.py:394-402) with `args` an `argparse.Namespace` and `module.`-prefixed keys (DDP).  torch >= 2.6 loads
    with `weights_only=True` by default, so the Namespace is allow-listed explicitly.  Like the reference's inference drivers
    (main_infer_narrator.py:107 `load_state_dict(strict=True)`), a key mismatch raises; only re-derivable buffers and a
    `temporal_embed` of a different length (handled by `inflate_positional_embeds`, eval_zeroshot.py:108-118) are tolerated."""
    if checkpoint is None:
        print("=> no checkpoint given and no network: random initialisation (reference would load OpenAI CLIP weights)")
        return
    if isinstance(checkpoint, str):
        import argparse
        with torch.serialization.safe_globals([argparse.Namespace]):
            sd = torch.load(checkpoint, map_location='cpu', weights_only=True)
    else:
        sd = checkpoint
    sd = sd.get('state_dict', sd)
    sd = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in sd.items()}
    res = model.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if not any(b in k for b in _BENIGN_MISSING)]
    unexpected = [k for k in res.unexpected_keys if not any(b in k for b in _BENIGN_MISSING)]
    if (missing or unexpected) and strict:
        raise RuntimeError("checkpoint does not match the model: missing %s, unexpected %s" % (missing[:8], unexpected[:8]))
    print("=> loaded checkpoint: %d tensors, %d benign buffer keys skipped" % (
        len(sd) - len(res.unexpected_keys), len(res.missing_keys) + len(res.unexpected_keys) - len(missing) - len(unexpected)))


def _build(vision_kwargs, vision_width, text_width, text_heads, text_layers, num_frames, timesformer_gated_xattn,
           drop_path_rate, timesformer_freeze_space, temperature_init, project_embed_dim, checkpoint, kwargs):
    vision_model = SpaceTimeTransformer(num_frames=num_frames, time_init='zeros', attention_style='frozen-in-time',
                                        ln_pre=True, act_layer=QuickGELU, is_tanh_gating=timesformer_gated_xattn,
                                        drop_path_rate=drop_path_rate, **vision_kwargs)
    vision_model.head = nn.Identity()
    vision_model.pre_logits = nn.Identity()
    vision_model.fc = nn.Identity()
    model = CLIP(embed_dim=project_embed_dim, vision_width=vision_width, vision_model=vision_model, context_length=77,
                 vocab_size=49408, transformer_width=text_width, transformer_heads=text_heads,
                 transformer_layers=text_layers, tempearture_init=temperature_init, **kwargs)
    _load_checkpoint(model, checkpoint)
    if timesformer_freeze_space:
        for n, p in vision_model.named_parameters():
            p.requires_grad = ('temporal_embed' in n or 'timeattn' in n or 'norm3' in n or n == 'cls_token'
                               or 'alpha_timeattn' in n)
    return model


def CLIP_OPENAI_TIMESFORMER_BASE(num_frames=4, timesformer_gated_xattn=False, drop_path_rate=0,
                                 timesformer_freeze_space=False, temperature_init=0.07, project_embed_dim=256,
                                 checkpoint=None, **kwargs):
    """models.py:316-371: TSF-B/16 (768/12/12) + CLIP-B text tower (512/8/12)."""
    return _build({}, 768, 512, 8, 12, num_frames, timesformer_gated_xattn, drop_path_rate, timesformer_freeze_space,
                  temperature_init, project_embed_dim, checkpoint, kwargs)


def CLIP_OPENAI_TIMESFORMER_LARGE(num_frames=4, timesformer_gated_xattn=False, drop_path_rate=0,
                                  timesformer_freeze_space=False, temperature_init=0.07, project_embed_dim=256,
                                  checkpoint=None, **kwargs):
    """models.py:374-431: TSF-L/14 224px (1024/24/16) + CLIP-L text tower (768/12/12)."""
    return _build(dict(img_size=224, patch_size=14, embed_dim=1024, depth=24, num_heads=16), 1024, 768, 12, 12,
                  num_frames, timesformer_gated_xattn, drop_path_rate, timesformer_freeze_space, temperature_init,
                  project_embed_dim, checkpoint, kwargs)


def CLIP_OPENAI_TIMESFORMER_LARGE_336PX(num_frames=4, timesformer_gated_xattn=False, drop_path_rate=0,
                                        timesformer_freeze_space=False, temperature_init=0.07, project_embed_dim=256,
                                        checkpoint=None, **kwargs):
    """models.py:434-491: TSF-L/14 336px."""
    return _build(dict(img_size=336, patch_size=14, embed_dim=1024, depth=24, num_heads=16), 1024, 768, 12, 12,
                  num_frames, timesformer_gated_xattn, drop_path_rate, timesformer_freeze_space, temperature_init,
                  project_embed_dim, checkpoint, kwargs)


# ------------------------------------------------------------------------------------------------ narrator factories
def _vclm(vision_kwargs, vision_width, gpt2_cfg, cross_attn_freq, heads, num_frames, gated_xattn, timesformer_gated_xattn,
          freeze_lm_vclm, freeze_visual_vclm, freeze_visual_vclm_temporal, checkpoint, kwargs):
    from types import SimpleNamespace
    from .gpt2_gated import GPT2LMHeadModel as GatedGPT2LMHeadModel, augment_gpt2_config
    from .narrator import VCLM_HF
    vision_model = SpaceTimeTransformer(num_frames=num_frames, time_init='zeros', attention_style='frozen-in-time', ln_pre=True,
                                        act_layer=QuickGELU, is_tanh_gating=timesformer_gated_xattn, **vision_kwargs)
    vision_model.head = nn.Identity()
    vision_model.pre_logits = nn.Identity()
    vision_model.fc = nn.Identity()
    config = SimpleNamespace(layer_norm_epsilon=1e-5, activation_function="gelu_new", n_positions=1024, vocab_size=50257, **gpt2_cfg)
    text_decoder = GatedGPT2LMHeadModel(augment_gpt2_config(config, cross_attn_freq=cross_attn_freq, gated_xattn=gated_xattn))
    if freeze_lm_vclm:
        text_decoder.freeze_lm_weights()
    if freeze_visual_vclm:
        vision_model.freeze_spatial_weights()
    if freeze_visual_vclm_temporal:
        vision_model.freeze_temporal_weights()
    model = VCLM_HF(vision_width=vision_width, vision_model=vision_model, text_width=gpt2_cfg["n_embd"],
                    text_decoder=text_decoder, num_img_queries=256, dim_head=64, heads=heads, **kwargs)
    _load_checkpoint(model, checkpoint)
    return model


def VCLM_OPENAI_TIMESFORMER_BASE_GPT2(gated_xattn=False, random_init_gpt2=False, freeze_lm_vclm=False,
                                      freeze_visual_vclm=False, freeze_visual_vclm_temporal=False, num_frames=4,
                                      timesformer_gated_xattn=False, checkpoint=None, **kwargs):
    """models.py:887-948: TSF-B/16 + GPT-2 (768/12/12), cross-attention in every layer, 12 pooling heads."""
    return _vclm({}, 768, dict(n_embd=768, n_layer=12, n_head=12), 1, 12, num_frames, gated_xattn, timesformer_gated_xattn,
                 freeze_lm_vclm, freeze_visual_vclm, freeze_visual_vclm_temporal, checkpoint, kwargs)


def VCLM_OPENAI_TIMESFORMER_BASE_GPT2_XL(gated_xattn=False, random_init_gpt2=False, freeze_lm_vclm=False,
                                         freeze_visual_vclm=False, freeze_visual_vclm_temporal=False, num_frames=4,
                                         timesformer_gated_xattn=False, checkpoint=None, **kwargs):
    """models.py:951-1009: TSF-B/16 + GPT-2 XL (1600/48/25), cross-attention every 2nd layer, 25 pooling heads."""
    return _vclm({}, 768, dict(n_embd=1600, n_layer=48, n_head=25), 2, 25, num_frames, gated_xattn, timesformer_gated_xattn,
                 freeze_lm_vclm, freeze_visual_vclm, freeze_visual_vclm_temporal, checkpoint, kwargs)


_TSF_L = dict(img_size=224, patch_size=14, embed_dim=1024, depth=24, num_heads=16)
_TSF_L_336 = dict(img_size=336, patch_size=14, embed_dim=1024, depth=24, num_heads=16)


def VCLM_OPENAI_TIMESFORMER_LARGE_GPT2_XL(gated_xattn=False, random_init_gpt2=False, freeze_lm_vclm=False,
                                          freeze_visual_vclm=False, freeze_visual_vclm_temporal=False, num_frames=4,
                                          timesformer_gated_xattn=False, checkpoint=None, **kwargs):
    """models.py:1012-1072 (the NARRATOR of BASELINE config 4): TSF-L/14 224px (256 patches per frame: key-tiled space
    attention, csrc/attention_big.cu) + GPT-2 XL, cross-attention every 2nd layer, 25 pooling heads."""
    return _vclm(dict(_TSF_L), 1024, dict(n_embd=1600, n_layer=48, n_head=25), 2, 25, num_frames, gated_xattn,
                 timesformer_gated_xattn, freeze_lm_vclm, freeze_visual_vclm, freeze_visual_vclm_temporal, checkpoint, kwargs)


def VCLM_OPENAI_TIMESFORMER_LARGE_GPT2(gated_xattn=False, random_init_gpt2=False, freeze_lm_vclm=False,
                                       freeze_visual_vclm=False, freeze_visual_vclm_temporal=False, num_frames=4,
                                       timesformer_gated_xattn=False, checkpoint=None, **kwargs):
    """models.py:1075-1135: TSF-L/14 224px + GPT-2 (768/12/12), cross-attention in every layer, 12 pooling heads."""
    return _vclm(dict(_TSF_L), 1024, dict(n_embd=768, n_layer=12, n_head=12), 1, 12, num_frames, gated_xattn,
                 timesformer_gated_xattn, freeze_lm_vclm, freeze_visual_vclm, freeze_visual_vclm_temporal, checkpoint, kwargs)


def VCLM_OPENAI_TIMESFORMER_LARGE_336PX_GPT2_XL(gated_xattn=False, random_init_gpt2=False, freeze_lm_vclm=False,
                                                freeze_visual_vclm=False, freeze_visual_vclm_temporal=False, num_frames=4,
                                                timesformer_gated_xattn=False, checkpoint=None, **kwargs):
    """models.py:1138-1198: TSF-L/14 336px (576 patches per frame) + GPT-2 XL, cross-attention every 3rd layer."""
    return _vclm(dict(_TSF_L_336), 1024, dict(n_embd=1600, n_layer=48, n_head=25), 3, 25, num_frames, gated_xattn,
                 timesformer_gated_xattn, freeze_lm_vclm, freeze_visual_vclm, freeze_visual_vclm_temporal, checkpoint, kwargs)


# ------------------------------------------------------------------------------------------------ CLIP_HF factories
def _build_hf(vision_kwargs, num_frames, timesformer_gated_xattn, drop_path_rate, timesformer_freeze_space, temperature_init,
              project_embed_dim, text_model, checkpoint, kwargs):
    vision_model = SpaceTimeTransformer(num_frames=num_frames, time_init='zeros', attention_style='frozen-in-time',
                                        ln_pre=True, act_layer=QuickGELU, is_tanh_gating=timesformer_gated_xattn,
                                        drop_path_rate=drop_path_rate, **vision_kwargs)
    vision_model.head = nn.Identity()
    vision_model.pre_logits = nn.Identity()
    vision_model.fc = nn.Identity()
    if text_model is None:
        from transformers import DistilBertConfig, DistilBertModel
        print("=> no network: DistilBERT-base is randomly initialised (reference loads 'distilbert-base-uncased')")
        text_model = DistilBertModel(DistilBertConfig())
    kwargs.pop('text_use_cls_token', None)   # models.py:533: DistilBERT has no pooler, the CLS token is always used
    model = CLIP_HF(embed_dim=project_embed_dim, vision_width=vision_model.embed_dim, vision_model=vision_model,
                    text_width=text_model.config.hidden_size, text_model=text_model, text_use_cls_token=True,
                    text_is_regressive=False, tempearture_init=temperature_init, **kwargs)
    _load_checkpoint(model, checkpoint)
    if timesformer_freeze_space:
        for n, p in vision_model.named_parameters():
            p.requires_grad = ('temporal_embed' in n or 'timeattn' in n or 'norm3' in n or n == 'cls_token'
                               or 'alpha_timeattn' in n)
    return model


def CLIP_OPENAI_TIMESFORMER_BASE_DISTILBERT_BASE(num_frames=4, timesformer_gated_xattn=False, drop_path_rate=0,
                                                 timesformer_freeze_space=False, temperature_init=0.07, project_embed_dim=256,
                                                 text_model=None, checkpoint=None, **kwargs):
    """models.py:494-545: TSF-B/16 + DistilBERT-base text tower (CLS token)."""
    return _build_hf({}, num_frames, timesformer_gated_xattn, drop_path_rate, timesformer_freeze_space, temperature_init,
                     project_embed_dim, text_model, checkpoint, kwargs)


def CLIP_OPENAI_TIMESFORMER_LARGE_DISTILBERT_BASE(num_frames=4, timesformer_gated_xattn=False, drop_path_rate=0,
                                                  timesformer_freeze_space=False, temperature_init=0.07, project_embed_dim=256,
                                                  text_model=None, checkpoint=None, **kwargs):
    """models.py:548-601: TSF-L/14 224px + DistilBERT-base."""
    return _build_hf(dict(_TSF_L), num_frames, timesformer_gated_xattn, drop_path_rate, timesformer_freeze_space,
                     temperature_init, project_embed_dim, text_model, checkpoint, kwargs)


def CLIP_OPENAI_TIMESFORMER_LARGE_336PX_DISTILBERT_BASE(num_frames=4, timesformer_gated_xattn=False, drop_path_rate=0,
                                                        timesformer_freeze_space=False, temperature_init=0.07,
                                                        project_embed_dim=256, text_model=None, checkpoint=None, **kwargs):
    """models.py:604-657: TSF-L/14 336px + DistilBERT-base (the TSF-L@HR recipe, docs/PRETRAIN.md:24-36)."""
    return _build_hf(dict(_TSF_L_336), num_frames, timesformer_gated_xattn, drop_path_rate, timesformer_freeze_space,
                     temperature_init, project_embed_dim, text_model, checkpoint, kwargs)
