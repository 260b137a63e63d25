"""Checkpoint plumbing of lavila/models/utils.py for the B200 mirror: temporal-embedding inflation when a checkpoint was
trained with a different number of frames (utils.py:13-59, used by main_pretrain.py / eval_zeroshot.py / the narrator
scripts when loading released 4-frame checkpoints into 16-frame models), OpenAI-CLIP -> TimeSformer key remapping
(utils.py:74-105) and the dotted-attribute helpers (utils.py:62-71).  Pure host-side tensor bookkeeping: no kernels."""
import functools
import re
from collections import OrderedDict

import torch
import torch.nn.functional as F

_TEMPORAL, _POS = 'visual.temporal_embed', 'visual.pos_embed'


def _resample_frames(embed, frames, how):
    """embed [1, T0, D] -> [1, frames, D] for frames > T0."""
    if how == 'zeros':
        out = torch.zeros(embed.shape[0], frames, embed.shape[2])
        out[:, :embed.shape[1]] = embed
        return out
    if how in ('interp', 'bilinear'):
        # treated as a 1-channel image of size T0 x D resized to frames x D (utils.py:39-46)
        return F.interpolate(embed.unsqueeze(0), (frames, embed.shape[2]),
                             mode='bilinear' if how == 'bilinear' else 'nearest').squeeze(0)
    raise NotImplementedError(how)


def inflate_positional_embeds(current_model_state_dict, new_state_dict, num_frames=4, load_temporal_fix='bilinear'):
    """Adapt `new_state_dict` (a checkpoint) to a model built for `num_frames`: more frames in the checkpoint -> keep
    the first `num_frames` temporal embeddings; fewer -> fill by zeros / nearest / bilinear interpolation.  A different
    number of spatial patches is refused, like the reference (utils.py:50-57).  Returns the (modified) checkpoint dict."""
    have = set(current_model_state_dict.keys())
    if _TEMPORAL in new_state_dict and _TEMPORAL in have:
        loaded = new_state_dict[_TEMPORAL]
        t_ckpt = loaded.shape[1]
        if t_ckpt > num_frames:
            print('### loaded SpaceTimeTransformer model has MORE frames than current...'
                  '### loading weights, filling in the extras via {}'.format(load_temporal_fix))
            new_state_dict[_TEMPORAL] = loaded[:, :num_frames, :]
        elif t_ckpt < num_frames:
            print('### loaded SpaceTimeTransformer model has FEWER frames than current...'
                  '### loading weights, filling in the extras via {}'.format(load_temporal_fix))
            new_state_dict[_TEMPORAL] = _resample_frames(loaded, num_frames, load_temporal_fix)
    if _POS in new_state_dict and _POS in have:
        if new_state_dict[_POS].shape[1] != current_model_state_dict[_POS].shape[1]:
            raise NotImplementedError(
                'Loading models with different spatial resolution / patch number not yet implemented, sorry.')
    return new_state_dict


def rgetattr(obj, attr, *args):
    """getattr over a dotted path (utils.py:67-70)."""
    return functools.reduce(lambda o, name: getattr(o, name, *args), attr.split('.'), obj)


def rsetattr(obj, attr, val):
    """setattr over a dotted path (utils.py:62-64)."""
    head, _, leaf = attr.rpartition('.')
    return setattr(rgetattr(obj, head) if head else obj, leaf, val)


_STEM_KEYS = {
    "class_embedding": "cls_token", "positional_embedding": "pos_embed", "conv1.weight": "patch_embed.proj.weight",
    "ln_pre.weight": "ln_pre.weight", "ln_pre.bias": "ln_pre.bias", "ln_post.weight": "norm.weight",
    "ln_post.bias": "norm.bias",
}
_BLOCK_KEYS = {
    "attn.in_proj_weight": "attn.qkv.weight", "attn.in_proj_bias": "attn.qkv.bias", "attn.out_proj.weight": "attn.proj.weight",
    "attn.out_proj.bias": "attn.proj.bias", "ln_1.weight": "norm1.weight", "ln_1.bias": "norm1.bias",
    "mlp.c_fc.weight": "mlp.fc1.weight", "mlp.c_fc.bias": "mlp.fc1.bias", "mlp.c_proj.weight": "mlp.fc2.weight",
    "mlp.c_proj.bias": "mlp.fc2.bias", "ln_2.weight": "norm2.weight", "ln_2.bias": "norm2.bias",
}
_RESBLOCK = re.compile(r"^transformer\.resblocks\.(\d+)\.(.+)$")


def remap_keys(clip_state_dict, transformer_layers=12):
    """OpenAI CLIP `visual.*` state dict -> SpaceTimeTransformer names (utils.py:74-105).  `proj` is skipped (loaded
    separately as image_projection); class / positional embeddings gain their leading singleton dimensions (and, like the
    reference, are updated in the input dict too).  A key outside the table raises KeyError, as the reference does."""
    out = OrderedDict()
    for key in clip_state_dict:
        if key == 'proj':
            continue
        m = _RESBLOCK.match(key)
        if m is not None and int(m.group(1)) < transformer_layers and m.group(2) in _BLOCK_KEYS:
            new_key = "blocks.{}.{}".format(int(m.group(1)), _BLOCK_KEYS[m.group(2)])
        else:
            new_key = _STEM_KEYS[key]
        if key == "class_embedding":
            clip_state_dict[key] = clip_state_dict[key].unsqueeze(0).unsqueeze(0)
        elif key == "positional_embedding":
            clip_state_dict[key] = clip_state_dict[key].unsqueeze(0)
        out[new_key] = clip_state_dict[key]
    return out
