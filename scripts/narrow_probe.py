"""Probe of BASELINE.json configs[4] on ONE GPU: does the n^3 narrow-band hierarchy build, how long do setup and a solve take?
  python scripts/narrow_probe.py 1024"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from geometricmultigridpressuresolver_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = api.Context(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
t0 = time.perf_counter()
out = bench.measure_narrow(torch, api, ctx, None, 0, 1, 0, n, flush)
out["wall_s"] = time.perf_counter() - t0
out["max_device_GB"] = torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9
print(json.dumps(out))
