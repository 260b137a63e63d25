"""The input-pipeline leg of bench.py alone (lv_clip_transform at BASELINE config 2's batch); prints its JSON object."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(bench.input_pipeline_leg(torch.device("cuda", 0))))
