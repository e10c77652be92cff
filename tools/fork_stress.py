"""Back-to-back calls (hundreds, no host synchronisation in between) on batches that hold plain tiles, pair tiles or both (the
pair launch is forked to a side stream): forward API (per-jet times) and / or short samplers; every result is compared with the
first one (the calls are deterministic).  usage: fork_stress.py <model> <plain|pair|mixed> <forward|generate|both> [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.mmf import time_grid
from mmf_b200.param_spec import make_config
model, kind, api = sys.argv[1], sys.argv[2], sys.argv[3]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 400
dev = torch.device("cuda:0")
cfg = make_config(model, num_timesteps=int(os.environ.get("STRESS_N", "3")))
nm = _abi.NativeModel(cfg, synthetic.make_state_dict(cfg, "wide", 0), dev)
g = torch.Generator().manual_seed(1)
ns = {"plain": [1, 9, 40, 77, 128, 33, 64, 100, 127, 5], "pair": [150, 140, 129, 133], "mixed": [1, 9, 40, 77, 128, 150, 140, 129, 33, 64]}[kind]
n = torch.tensor(ns)
B = len(n)
mask = synthetic.prefix_masks(n, 150)
x0 = (torch.randn(B, 150, 3, generator=g) * mask).to(dev)
k0 = (torch.randint(1, 9, (B, 150, 1), generator=g) * mask).to(dev)
t = torch.rand(B, generator=g).to(dev)
md = mask.to(dev)
ts, dt = time_grid(cfg)
ref = None
for it in range(iters):
    cur = []
    if api in ("forward", "both"):
        cur += list(nm.forward(x0, k0, md, t))
    if api in ("generate", "both"):
        cur += list(nm.generate(x0, k0, None, ts, dt, _abi.step_options(cfg, seed=2), n_per_jet=n.to(torch.int32))[:2])
    if ref is None:
        ref = [c.clone() for c in cur]
    elif it % 50 == 0 or it == iters - 1:
        assert all(torch.equal(a, b) for a, b in zip(cur, ref)), it
torch.cuda.synchronize()
nm.status()
print(model, kind, api, "ok", iters, flush=True)
