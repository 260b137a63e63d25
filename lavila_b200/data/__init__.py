"""GPU input pipeline: what the reference's DataLoader workers do to decoded frames, as one kernel per batch.
Decoding itself (decord on the CPU in the reference, lavila/data/datasets.py:25-75) is not part of this package."""
from .video_transforms import GpuClipTransform, Permute, center_crop_offsets, random_resized_crop_params, resize_output_size, transforms_for_model  # noqa: F401
from .datasets import get_frame_ids  # noqa: F401
