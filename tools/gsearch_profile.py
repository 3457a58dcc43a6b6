"""One call of the default global pivot finder at config-4 shape (bench.block_globalsearch), for ncu.
usage: ncu ... python tools/gsearch_profile.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tci_b200 as T  # noqa: E402

torch.cuda.set_device(0)
print(json.dumps(bench.block_globalsearch(T, T.default_context())))
