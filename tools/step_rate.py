"""Standalone step kernel, production mode: time per launch on a batch far larger than L2 (bench.py's step roofline)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
import bench
out = bench.step_kernel_roofline(bench.load_peaks(), torch.device("cuda:0"))
print(json.dumps(out))
