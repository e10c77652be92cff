"""Per-phase clock trace of the EPiC tile kernel (debugging aid): MMF_TRACE=<file> python tools/epic_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from mmf_b200.mmf import time_grid
cfg = make_config("EPiC", num_timesteps=4)
sd = synthetic.make_state_dict(cfg, "wide", 0)
nm = _abi.NativeModel(cfg, sd, torch.device("cuda:0"))
src = synthetic.source_state(256).to("cuda:0")
ts, dt = time_grid(cfg)
for _ in range(2):
    nm.generate(src.continuous, None, src.mask, ts, dt, None)
torch.cuda.synchronize()
print(open(os.environ["MMF_TRACE"]).read())
