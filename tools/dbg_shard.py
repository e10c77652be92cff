import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from oracle import mmf_oracle as orc
DEV = torch.device("cuda:0")
for model in ("FusedParticleFormer", "ParticleFormer"):
  for nt in (1, 2, 8):
    cfg = make_config(model, num_timesteps=nt)
    sd = synthetic.make_state_dict(cfg, "wide", 0)
    nm = _abi.NativeModel(cfg, sd, DEV)
    src = synthetic.source_state(16, seed=55).to(DEV)
    ts, dt = orc.time_grid(cfg)
    x, k, _ = nm.generate(src.continuous, src.discrete, src.mask, ts, float(dt), _abi.step_options(cfg, seed=5, first_global_jet=32))
    parts = []
    for lo, hi in ((0, 6), (6, 16)):
        s = src[lo:hi]
        parts.append(nm.generate(s.continuous, s.discrete, s.mask, ts, float(dt), _abi.step_options(cfg, seed=5, first_global_jet=32 + lo)))
    torch.cuda.synchronize()
    kc, xc = torch.cat([p[1] for p in parts]), torch.cat([p[0] for p in parts])
    real = src.mask.bool().squeeze(-1)
    a, b = xc[real].float(), x[real].float()
    print(model, nt, "k agree", float((kc[real] == k[real]).float().mean()), "x rel", float((a - b).norm() / b.norm()), "max abs", float((a - b).abs().max()),
          "n", [int(m.sum()) for m in src.mask.squeeze(-1)])
    # forward API: velocity at t for the same jets, one call vs shards
    t = torch.full((16,), 0.3, device=DEV)
    va, la = nm.forward(src.continuous, src.discrete, src.mask, t)
    vs, ls = [], []
    for lo, hi in ((0, 6), (6, 16)):
        s = src[lo:hi]
        v_, l_ = nm.forward(s.continuous, s.discrete, s.mask, t[lo:hi])
        vs.append(v_); ls.append(l_)
    vb, lb = torch.cat(vs), torch.cat(ls)
    print("   forward: vt rel", float((va[real] - vb[real]).norm() / va[real].norm()), "logits rel", float((la[real] - lb[real]).norm() / la[real].norm()))
