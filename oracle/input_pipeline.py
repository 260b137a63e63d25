"""TEST INFRASTRUCTURE: plain restatement of the reference's per-sample clip transform chain, for the GPU input pipeline
(`lv_clip_transform`, lavila_b200/data/).  Never imported by the product.

Follows main_pretrain.py:263-281: `Permute([3, 0, 1, 2])` (lavila/data/video_transforms.py:15-32), then
train `transforms.RandomResizedCrop(crop, scale=(0.5, 1.0))` / val `transforms.Resize(crop)` + `transforms.CenterCrop(crop)`,
then `transforms_video.NormalizeVideo(mean, std)`.  torchvision is a third-party dependency of the reference (0.11.2 pinned in
requirements.txt:3, 0.26 installed here); its tensor path is `crop` + `torch.nn.functional.interpolate(mode="bilinear",
align_corners=False, antialias=...)` (antialias False in 0.11.2, True by default from 0.17) + `(clip - mean) / std`.

Pinned: tests/golden/make_golden_input_pipeline.py runs the reference's own `Permute` with the installed torchvision transforms
(seeded) and stores inputs, boxes and outputs in tests/golden/input_pipeline.pt; tests/test_oracle_input_pipeline.py checks
this file against them, and `get_frame_ids` against a recorded run of lavila/data/datasets.py:78-90."""
import numpy as np
import torch
import torch.nn.functional as F


def normalize_video(clip_cthw, mean, std):
    m = torch.as_tensor(mean, dtype=torch.float32).view(-1, 1, 1, 1)
    s = torch.as_tensor(std, dtype=torch.float32).view(-1, 1, 1, 1)
    return (clip_cthw - m) / s


def train_transform(frames_thwc, box, size, mean, std, antialias=False):
    """box = (i, j, h, w) as drawn by RandomResizedCrop.get_params."""
    i, j, h, w = box
    clip = frames_thwc.float().permute(3, 0, 1, 2)                              # Permute: T H W C -> C T H W
    clip = clip[..., i:i + h, j:j + w]                                          # F.crop
    clip = F.interpolate(clip, size=(size, size), mode="bilinear", align_corners=False, antialias=antialias)   # C is the batch dim
    return normalize_video(clip, mean, std)


def resized_size(h, w, size):
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def val_transform(frames_thwc, size, mean, std, antialias=False):
    clip = frames_thwc.float().permute(3, 0, 1, 2)
    H, W = clip.shape[-2:]
    rh, rw = resized_size(H, W, size)
    clip = F.interpolate(clip, size=(rh, rw), mode="bilinear", align_corners=False, antialias=antialias)
    top, left = int(round((rh - size) / 2.0)), int(round((rw - size) / 2.0))
    return normalize_video(clip[..., top:top + size, left:left + size], mean, std)


def get_frame_ids(start_frame, end_frame, num_segments=32, jitter=True):
    """lavila/data/datasets.py:78-90."""
    seg_size = float(end_frame - start_frame - 1) / num_segments
    seq = []
    for i in range(num_segments):
        start = int(np.round(seg_size * i) + start_frame)
        end = min(int(np.round(seg_size * (i + 1)) + start_frame), end_frame)
        seq.append(np.random.randint(low=start, high=(end + 1)) if jitter else (start + end) // 2)
    return seq
