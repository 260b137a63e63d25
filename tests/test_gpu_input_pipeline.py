"""lv_clip_transform on the GPU (through GpuClipTransform -> ops -> the C ABI) against the oracle restatement of the reference's
transform chain (oracle/input_pipeline.py, pinned by tests/golden/input_pipeline.pt) and against those goldens directly.
Floating point: |diff| <= 2e-5 on normalised values (|v| <= ~2.3) for plain bilinear -- a 4-tap sum of 0..255 values in fp32,
evaluated in a different order than ATen's CPU kernel -- and <= 1e-4 with antialiasing (up to ~5 x 5 taps; ATen's CPU kernel runs
two separable passes with an fp32 intermediate, the kernel one 2-D sum: observed 4.5e-5)."""
import os

import pytest
import torch

from oracle import input_pipeline as O

pytestmark = pytest.mark.gpu
GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.pt"), weights_only=False)
TOL = 2e-5
TOL_AA = 1e-4


@pytest.mark.parametrize("src_dtype", [torch.uint8, torch.float32])
def test_golden_cases(src_dtype):
    from lavila_b200.data import GpuClipTransform
    for c in GOLD["train"]:
        tf = GpuClipTransform(c["crop"], "train", c["mean"], c["std"], antialias=c["antialias"])
        torch.manual_seed(c["seed"])
        out = tf([c["frames"].to(src_dtype)])
        assert tf.last_boxes[0][:4] == c["box"]
        assert float((out[0].cpu() - c["out"]).abs().max()) <= (TOL_AA if c["antialias"] else TOL), ("train", c["crop"], c["antialias"])
    for c in GOLD["val"]:
        tf = GpuClipTransform(c["crop"], "val", c["mean"], c["std"], antialias=c["antialias"])
        out = tf([c["frames"].to(src_dtype).cuda()])
        assert float((out[0].cpu() - c["out"]).abs().max()) <= (TOL_AA if c["antialias"] else TOL), ("val", c["crop"], c["antialias"])


@pytest.mark.parametrize("antialias", [False, True])
def test_ragged_batch_at_the_real_geometry(antialias):
    """Six clips of different source sizes in one launch (Ego4D chunks are 288 px on the short side, landscape or portrait),
    16 frames -> 224 x 224, seeded boxes; train and val."""
    from lavila_b200.data import GpuClipTransform, video_transforms as VT
    g = torch.Generator().manual_seed(3)
    sizes = [(288, 384), (288, 512), (384, 288), (240, 320), (288, 288), (360, 640)]
    clips = [torch.randint(0, 256, (16, h, w, 3), generator=g, dtype=torch.uint8) for h, w in sizes]
    tf = GpuClipTransform(224, "train", antialias=antialias)
    torch.manual_seed(9)
    out = tf(clips)
    assert out.shape == (6, 3, 16, 224, 224) and out.dtype == torch.float32
    for k, c in enumerate(clips):
        ref = O.train_transform(c, tf.last_boxes[k][:4], 224, VT.OPENAI_MEAN, VT.OPENAI_STD, antialias)
        assert float((out[k].cpu() - ref).abs().max()) <= (TOL_AA if antialias else TOL), k
    out = GpuClipTransform(224, "val", antialias=antialias)(clips)
    for k, c in enumerate(clips):
        ref = O.val_transform(c, 224, VT.OPENAI_MEAN, VT.OPENAI_STD, antialias)
        assert float((out[k].cpu() - ref).abs().max()) <= (TOL_AA if antialias else TOL), k


def test_identity_geometry_is_exact_and_frame_stride_is_honoured():
    """box = frame, resized size = frame size: every tap weight is exactly 1, so the result is (v - mean) / std bit for bit;
    the clips are strided views (every 2nd frame of a longer buffer)."""
    from lavila_b200.data import GpuClipTransform, video_transforms as VT
    g = torch.Generator().manual_seed(4)
    buf = torch.randint(0, 256, (3, 8, 64, 64, 3), generator=g, dtype=torch.uint8).cuda()
    clips = [buf[b, ::2] for b in range(3)]
    tf = GpuClipTransform(64, "val")
    out = tf(clips, boxes=[(0, 0, 64, 64, 64, 64, 0, 0)] * 3)
    m = torch.tensor(VT.OPENAI_MEAN, device="cuda").view(3, 1, 1, 1)
    s = torch.tensor(VT.OPENAI_STD, device="cuda").view(3, 1, 1, 1)
    for b in range(3):
        assert torch.equal(out[b], (clips[b].float().permute(3, 0, 1, 2) - m) / s)


def test_full_batch_properties():
    """BASELINE config 2's batch (64 clips x 16 frames, 288 x 384 sources -> 224^2): every clip of the batched launch equals the
    same clip transformed alone with the same box (bit exact), and the output feeds the model's input layout."""
    from lavila_b200.data import GpuClipTransform
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.randint(0, 256, (64, 16, 288, 384, 3), generator=g, dtype=torch.uint8, device="cuda")
    tf = GpuClipTransform(224, "train")
    torch.manual_seed(1)
    out = tf(src)
    assert out.shape == (64, 3, 16, 224, 224) and bool(torch.isfinite(out).all())
    boxes = list(tf.last_boxes)
    for b in (0, 17, 63):
        assert torch.equal(tf([src[b]], boxes=[boxes[b]])[0], out[b])
