"""Frame-index selection of the reference's video loaders (host logic, no decoding here)."""
import numpy as np


def get_frame_ids(start_frame, end_frame, num_segments=32, jitter=True):
    """One frame per segment of [start_frame, end_frame): the segment's middle, or a uniform draw from it with `jitter`
    (np.random, one `randint` per segment in order, like lavila/data/datasets.py:78-90)."""
    seg = float(end_frame - start_frame - 1) / num_segments
    ids = []
    for k in range(num_segments):
        lo = int(np.round(seg * k) + start_frame)
        hi = min(int(np.round(seg * (k + 1)) + start_frame), end_frame)
        ids.append(np.random.randint(low=lo, high=hi + 1) if jitter else (lo + hi) // 2)
    return ids
