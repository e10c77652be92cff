"""Loss curves of the device training step and of fp32 torch (autograd over the oracle + torch.optim.Adam + clip_grad_norm_) on the
same stream of synthetic batches with the same supplied draws (time, bridge noise, categorical uniforms).
usage: train_curve.py [model] [steps] [jets] [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import synthetic
from mmf_b200.mmf import MultiModalFlowBridge
from mmf_b200.param_spec import make_config
from oracle import mmf_oracle as orc

model = sys.argv[1] if len(sys.argv) > 1 else "ParticleFormer"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
out_path = sys.argv[4] if len(sys.argv) > 4 else None
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
cfg = make_config(model, lr=5e-4, sigma=1e-3)
sd = synthetic.make_state_dict(cfg, "default", 0)            # the reference's own initialisation scale
bridge = MultiModalFlowBridge(cfg)
bridge.model.load_state_dict(sd)
sd_loss = {k: v.detach().clone() for k, v in bridge.loss_combine.state_dict().items()}
bridge = bridge.to(dev)
eng = bridge.configure_training(lr=cfg.lr)
sdg = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in sd.items()}
slg = {k: torch.nn.Parameter(v.to(dev).clone()) for k, v in sd_loss.items()}
params = list(sdg.values()) + list(slg.values())
opt = torch.optim.Adam(params, lr=cfg.lr)
# a small "dataset": 8 fixed batches cycled (so that the loss can go down), fresh draws every step
data = [synthetic.training_batch(B, seed=100 + i) for i in range(8)]
gen = torch.Generator().manual_seed(7)
mine, theirs = [], []
for step in range(steps):
    b = data[step % len(data)]
    D = b.target.mask.shape[1]
    t = cfg.time_eps + (1 - cfg.time_eps) * torch.rand(B, generator=gen)
    z, u = torch.randn(B, D, 3, generator=gen), torch.rand(B, D, generator=gen)
    out5 = eng.train_step(b, time=t, z=z, u=u)
    mine.append(float(out5[0]))
    opt.zero_grad()
    d = lambda x: x.to(dev)
    loss = orc.training_loss(sdg, slg, cfg, d(b.source.continuous), d(b.source.discrete), d(b.target.continuous), d(b.target.discrete),
                             d(b.target.mask), d(t), d(z), d(u))[0]
    loss.backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()
    theirs.append(float(loss))
w = max(1, steps // 10)
avg = lambda v, i: sum(v[i:i + w]) / len(v[i:i + w])
res = {"model": model, "steps": steps, "jets_per_step": B, "lr": cfg.lr, "window": w,
       "device_path": {"first": avg(mine, 0), "last": avg(mine, steps - w)}, "fp32_torch": {"first": avg(theirs, 0), "last": avg(theirs, steps - w)},
       "max_rel_diff_of_window_means": max(abs(avg(mine, i) - avg(theirs, i)) / abs(avg(theirs, i)) for i in range(0, steps - w + 1, w)),
       "curve_device_path": [round(v, 4) for v in mine], "curve_fp32_torch": [round(v, 4) for v in theirs]}
print(json.dumps({k: v for k, v in res.items() if not k.startswith("curve")}))
if out_path:
    json.dump(res, open(out_path, "w"))
