"""Input pipeline without a GPU: the oracle (oracle/input_pipeline.py) against goldens of the reference's transform chain
(tests/golden/make_golden_input_pipeline.py), and the product's host logic (crop-box sampling with torchvision's RNG calls, frame-id
selection, descriptor geometry) with the kernel entry point replaced by its test double."""
import os

import numpy as np
import pytest
import torch

from oracle import input_pipeline as O

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.pt"), weights_only=False)


def test_oracle_train_transform_equals_reference_chain():
    for c in GOLD["train"]:
        out = O.train_transform(c["frames"], c["box"], c["crop"], c["mean"], c["std"], c["antialias"])
        assert out.shape == c["out"].shape
        assert float((out - c["out"]).abs().max()) <= 1e-5, (c["crop"], c["antialias"])


def test_oracle_val_transform_equals_reference_chain():
    for c in GOLD["val"]:
        out = O.val_transform(c["frames"], c["crop"], c["mean"], c["std"], c["antialias"])
        assert out.shape == c["out"].shape
        assert float((out - c["out"]).abs().max()) <= 1e-5


@pytest.mark.parametrize("which", ["oracle", "product"])
def test_get_frame_ids_equals_reference(which):
    from lavila_b200.data import get_frame_ids
    fn = O.get_frame_ids if which == "oracle" else get_frame_ids
    for c in GOLD["frame_ids"]:
        np.random.seed(c["np_seed"])
        assert [int(v) for v in fn(*c["args"])] == c["ids"], c["args"]


def test_crop_box_sampling_consumes_the_rng_like_torchvision():
    from lavila_b200.data import random_resized_crop_params
    for c in GOLD["train"]:
        T, H, W, _ = c["frames"].shape
        torch.manual_seed(c["seed"])
        assert random_resized_crop_params(H, W, (0.5, 1.0)) == c["box"]
    tv = pytest.importorskip("torchvision.transforms")
    for seed, (H, W, scale, ratio) in enumerate([(288, 384, (0.5, 1.0), (3 / 4, 4 / 3)), (240, 320, (0.08, 1.0), (3 / 4, 4 / 3)),
                                                 (50, 400, (0.9, 1.0), (0.9, 1.1)), (400, 50, (0.9, 1.0), (0.9, 1.1))]):
        torch.manual_seed(seed)
        ref = tv.RandomResizedCrop.get_params(torch.empty(3, H, W), scale, ratio)
        after_ref = torch.rand(1)
        torch.manual_seed(seed)
        got = random_resized_crop_params(H, W, scale, ratio)
        assert tuple(int(v) for v in ref) == got and torch.equal(after_ref, torch.rand(1))       # incl. the central fallback


def test_host_transform_on_double_equals_reference_chain(monkeypatch):
    """GpuClipTransform end to end on the CPU (descriptor table, box sampling, val geometry), kernel replaced by its double."""
    from tests import ops_doubles
    ops_doubles.install(monkeypatch)
    from lavila_b200.data import GpuClipTransform
    for c in GOLD["train"]:
        tf = GpuClipTransform(c["crop"], "train", c["mean"], c["std"], antialias=c["antialias"], device="cpu")
        torch.manual_seed(c["seed"])
        out = tf([c["frames"]])
        assert tf.last_boxes[0][:4] == c["box"]
        assert float((out[0] - c["out"]).abs().max()) <= 1e-5
    for c in GOLD["val"]:
        tf = GpuClipTransform(c["crop"], "val", c["mean"], c["std"], antialias=c["antialias"], device="cpu")
        out = tf(c["frames"].float().unsqueeze(0))
        assert float((out[0] - c["out"]).abs().max()) <= 1e-5


def test_transforms_for_model_follows_the_driver():
    from lavila_b200.data import transforms_for_model, video_transforms as VT
    t = transforms_for_model("CLIP_OPENAI_TIMESFORMER_LARGE_336PX", True, device="cpu")
    assert (t.crop_size, t.mode, t.mean, t.scale) == (336, "train", VT.OPENAI_MEAN, (0.5, 1.0))
    t = transforms_for_model("CLIP_TIMESFORMER_BASE", False, device="cpu")
    assert (t.crop_size, t.mode, t.mean, t.std) == (224, "val", VT.IMAGENET_MEAN, VT.IMAGENET_STD)
    with pytest.raises(ValueError):
        t(torch.zeros(1, 2, 100, 150, 3), boxes=[(0, 0, 100, 150, 100, 150, 0, 0)])                # window larger than the resized image


def test_product_refuses_the_cpu():
    from lavila_b200 import _lib as L
    from lavila_b200.data import GpuClipTransform
    with pytest.raises(L.LavilaB200Error):
        GpuClipTransform(32, "val", device="cpu")(torch.zeros(1, 2, 40, 40, 3))
