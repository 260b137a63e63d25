"""Golden vectors for the input pipeline: the reference's transform chain (main_pretrain.py:263-281) executed here with the
reference's own `Permute` (lavila/data/video_transforms.py) and the installed torchvision, plus `get_frame_ids` of
lavila/data/datasets.py run as is.

    python tests/golden/make_golden_input_pipeline.py      # writes tests/golden/input_pipeline.pt

torchvision 0.11.2 (the reference's pin) resizes tensors without antialiasing; the installed 0.26 takes `antialias` explicitly,
so both settings are recorded: `antialias=False` is the reference's pinned behaviour."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_shim  # noqa: E402

OPENAI = ([108.3272985, 116.7460125, 104.09373615000001], [68.5005327, 66.6321579, 70.32316305])
IMAGENET = ([123.675, 116.28, 103.53], [58.395, 57.12, 57.375])


def main():
    assert reference_shim.install()
    import torchvision
    from torchvision import transforms
    from torchvision.transforms import _transforms_video as transforms_video
    from lavila.data.video_transforms import Permute
    from lavila.data import datasets as RD
    gold = {"torchvision": torchvision.__version__, "train": [], "val": [], "frame_ids": []}
    g = torch.Generator().manual_seed(11)
    # (T, H, W, crop): a 288-short-side Ego4D-like frame, a portrait one, an odd-sized one, a frame smaller than the crop (up-sampling)
    for case, (T, H, W, crop) in enumerate([(4, 72, 96, 56), (3, 90, 60, 48), (2, 37, 53, 32), (2, 20, 28, 32)]):
        frames = torch.randint(0, 256, (T, H, W, 3), generator=g, dtype=torch.uint8)
        for aa in (False, True):
            for stats_name, (mean, std) in (("openai", OPENAI), ("imagenet", IMAGENET)):
                rrc = transforms.RandomResizedCrop(crop, scale=(0.5, 1.0), antialias=aa)
                torch.manual_seed(100 + case)
                box = rrc.get_params(torch.empty(3, T, H, W), rrc.scale, rrc.ratio)
                tf = transforms.Compose([Permute([3, 0, 1, 2]), rrc, transforms_video.NormalizeVideo(mean=mean, std=std)])
                torch.manual_seed(100 + case)
                out = tf(frames.float())                     # video_loader hands over fp32 frames (datasets.py:74-75)
                gold["train"].append({"frames": frames, "crop": crop, "antialias": aa, "stats": stats_name, "mean": mean,
                                      "std": std, "seed": 100 + case, "box": tuple(int(v) for v in box), "out": out})
                if min(H, W) >= 8:
                    vt = transforms.Compose([Permute([3, 0, 1, 2]), transforms.Resize(crop, antialias=aa),
                                             transforms.CenterCrop(crop), transforms_video.NormalizeVideo(mean=mean, std=std)])
                    gold["val"].append({"frames": frames, "crop": crop, "antialias": aa, "stats": stats_name, "mean": mean,
                                        "std": std, "out": vt(frames.float())})
    for (s, e, n, jit) in [(0, 300, 32, False), (17, 140, 16, False), (5, 9, 4, False), (0, 300, 32, True), (40, 77, 16, True),
                           (3, 20, 16, True)]:
        np.random.seed(1234)
        gold["frame_ids"].append({"args": (s, e, n, jit), "np_seed": 1234, "ids": [int(v) for v in RD.get_frame_ids(s, e, n, jit)]})
    path = os.path.join(ROOT, "tests", "golden", "input_pipeline.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB", len(gold["train"]), len(gold["val"]), len(gold["frame_ids"]))


if __name__ == "__main__":
    main()
