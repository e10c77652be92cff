"""Long run of the pair tiles (jets of more than 128 particles): FusedParticleFormer / ParticleFormer, 1000 timesteps, a batch with
many such jets next to small ones - 44 000 K / V exchanges per CTA pair; checks termination, finiteness and repeatability."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from mmf_b200.mmf import time_grid
dev = torch.device("cuda:0")
for model in ("FusedParticleFormer", "ParticleFormer"):
    cfg = make_config(model, num_timesteps=1000, temperature=1.2)
    nm = _abi.NativeModel(cfg, synthetic.make_state_dict(cfg, "wide", 0), dev)
    n = torch.cat([torch.randint(129, 151, (40,)), torch.randint(1, 129, (88,))])
    mask = synthetic.prefix_masks(n, 150)
    g = torch.Generator().manual_seed(4)
    x0 = (torch.randn(128, 150, 3, generator=g) * mask).to(dev)
    k0 = (torch.randint(1, 9, (128, 150, 1), generator=g) * mask).to(dev)
    ts, dt = time_grid(cfg)
    outs = []
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x, k, _ = nm.generate(x0, k0, mask.to(dev), ts, dt, _abi.step_options(cfg, seed=5), n_per_jet=n.to(torch.int32))
        e1.record()
        nm.status()
        outs.append((x.clone(), k.clone(), e0.elapsed_time(e1)))
    print(json.dumps({"model": model, "jets": 128, "jets_over_128": 40, "timesteps": 1000, "ms": round(outs[1][2], 1),
                      "finite": bool(torch.isfinite(outs[0][0]).all()), "repeatable": bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]))}), flush=True)
    nm.close()
