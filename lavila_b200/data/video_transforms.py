"""The clip transforms of main_pretrain.py:263-281 evaluated on the GPU for a whole batch (lv_clip_transform).

Reference chain per sample, on a DataLoader worker: `Permute([3, 0, 1, 2])` (lavila/data/video_transforms.py:15-32), then
train `transforms.RandomResizedCrop(crop, scale=(0.5, 1.0))` / val `transforms.Resize(crop)` + `transforms.CenterCrop(crop)`,
then `NormalizeVideo(mean, std)`.  The reference pins torchvision 0.11.2 (requirements.txt), whose tensor resize is plain bilinear
(`interpolate(mode="bilinear", align_corners=False)`): `antialias=False` is therefore the default here; `antialias=True` gives what
torchvision >= 0.17 does by default.  Crop boxes are drawn on the host with the same RNG calls as torchvision's `get_params`
(third-party dependency, algorithm restated from its published source), so a seeded run picks the same boxes."""
import math

import torch
import torch.nn as nn

from .. import ops

OPENAI_MEAN, OPENAI_STD = (108.3272985, 116.7460125, 104.09373615000001), (68.5005327, 66.6321579, 70.32316305)
IMAGENET_MEAN, IMAGENET_STD = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)


class Permute(nn.Module):
    """`frames.permute(ordering)` as a module (lavila/data/video_transforms.py:15-32); kept so that transform lists written for the
    reference still compose.  `GpuClipTransform` takes the decoder's T x H x W x C frames directly and needs no Permute."""

    def __init__(self, ordering):
        super().__init__()
        self.ordering = ordering

    def forward(self, frames):
        return frames.permute(self.ordering)


def random_resized_crop_params(height, width, scale=(0.5, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0)):
    """(i, j, h, w) as `torchvision.transforms.RandomResizedCrop.get_params`: up to 10 draws of (area fraction ~ U(scale),
    log aspect ~ U(log ratio)), position by two `torch.randint`; central fallback.  Same calls on the global torch RNG."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1]).item()
        aspect = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1])).item()
        w = int(round(math.sqrt(target_area * aspect)))
        h = int(round(math.sqrt(target_area / aspect)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,)).item()
            j = torch.randint(0, width - w + 1, size=(1,)).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def resize_output_size(height, width, size):
    """`transforms.Resize(int)`: the short side becomes `size`, the long side int(size * long / short)."""
    short, long = (width, height) if width <= height else (height, width)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if width <= height else (new_short, new_long)     # (new_h, new_w)


def center_crop_offsets(height, width, crop):
    """`transforms.CenterCrop`: top-left corner int(round((size - crop) / 2))."""
    return int(round((height - crop) / 2.0)), int(round((width - crop) / 2.0))


class GpuClipTransform:
    """clips (a list of T x H x W x 3 frame tensors, uint8 or fp32, on the host or the device; or one B x T x H x W x 3 tensor)
    -> fp32 [B, 3, T, crop, crop] on `device`, ready for `model(image, text)`.  mode "train": RandomResizedCrop + normalise
    (main_pretrain.py:263-272); mode "val": Resize + CenterCrop + normalise (:274-281)."""

    def __init__(self, crop_size=224, mode="train", mean=OPENAI_MEAN, std=OPENAI_STD, scale=(0.5, 1.0),
                 ratio=(3.0 / 4.0, 4.0 / 3.0), antialias=False, device="cuda"):
        if mode not in ("train", "val"):
            raise ValueError(f"mode {mode!r}")
        self.crop_size, self.mode, self.scale, self.ratio, self.antialias = int(crop_size), mode, scale, ratio, bool(antialias)
        self.mean, self.std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self.device = torch.device(device)
        self.last_boxes = None

    def geometry(self, height, width):
        """The 8 integers of one clip's descriptor after (H, W): box i, j, h, w, resized size, output-window offset."""
        S = self.crop_size
        if self.mode == "train":
            i, j, h, w = random_resized_crop_params(height, width, self.scale, self.ratio)
            return i, j, h, w, S, S, 0, 0
        rh, rw = resize_output_size(height, width, S)
        oy, ox = center_crop_offsets(rh, rw, S)
        return 0, 0, height, width, rh, rw, oy, ox

    def __call__(self, clips, boxes=None):
        return self.run(*self.prepare(clips, boxes))

    def run(self, desc, sources, frames):
        """The launch alone: descriptor table (device) + the frame tensors it points at -> the normalised batch."""
        out = torch.empty(len(sources), 3, frames, self.crop_size, self.crop_size, dtype=torch.float32, device=self.device)
        ops.clip_transform(desc, sources, frames, self.antialias, self.mean, self.std, out)
        return out

    def prepare(self, clips, boxes=None):
        """Host side of a batch: draw / check the geometry of every clip, move the frames to the device, upload the table."""
        if torch.is_tensor(clips):
            if clips.dim() != 5:
                raise ValueError("expected B x T x H x W x 3")
            clips = list(clips.unbind(0))
        if not clips:
            raise ValueError("empty batch")
        T = clips[0].shape[0]
        dt = clips[0].dtype
        if dt not in (torch.uint8, torch.float32):
            raise TypeError(f"frames must be uint8 or float32, got {dt}")
        rows, keep = [], []
        for k, c in enumerate(clips):
            if c.dim() != 4 or c.shape[3] != 3 or c.shape[0] != T or c.dtype != dt:
                raise ValueError("every clip must be T x H x W x 3 with the same T and dtype")
            H, W = int(c.shape[1]), int(c.shape[2])
            g = tuple(int(v) for v in boxes[k]) if boxes is not None else self.geometry(H, W)
            i, j, h, w, rh, rw, oy, ox = g
            if not (0 <= i and 0 <= j and h >= 1 and w >= 1 and i + h <= H and j + w <= W and oy >= 0 and ox >= 0
                    and oy + self.crop_size <= rh and ox + self.crop_size <= rw):
                raise ValueError(f"clip {k}: geometry {g} does not fit a {H} x {W} frame / crop {self.crop_size}")
            d = c.to(self.device, non_blocking=True)
            if d.stride(3) != 1 or d.stride(2) != 3 or d.stride(1) != 3 * W:
                d = d.contiguous()
            keep.append(d)
            rows.append((d.data_ptr(), H, W, i, j, h, w, rh, rw, oy, ox, d.stride(0)))
        self.last_boxes = [r[3:11] for r in rows]
        desc = torch.tensor(rows, dtype=torch.int64)
        if self.device.type == "cuda":
            desc = desc.pin_memory()
        return desc.to(self.device, non_blocking=True), keep, T


def transforms_for_model(model_name, is_training, device="cuda", antialias=False):
    """The transform main_pretrain.py builds for `args.model` (:262-281): 336 px crops for the 336PX models, the OpenAI CLIP
    statistics for the OPENAI families and the ImageNet ones otherwise."""
    crop = 336 if "336PX" in model_name else 224
    mean, std = (OPENAI_MEAN, OPENAI_STD) if "OPENAI" in model_name else (IMAGENET_MEAN, IMAGENET_STD)
    return GpuClipTransform(crop, "train" if is_training else "val", mean, std, antialias=antialias, device=device)
