"""Per-stage wall time of tools/pipeline_rate.py's loop (synchronised after every stage)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from mmf_b200.mmf import time_grid
model = sys.argv[1] if len(sys.argv) > 1 else "FusedParticleFormer"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda:0")
cfg = make_config(model, num_timesteps=100)
nm = _abi.NativeModel(cfg, synthetic.make_state_dict(cfg, "wide", 0), dev)
ts, dt = time_grid(cfg)
D, V = cfg.max_num_particles, cfg.vocab_size
n = np.clip(np.round(55 + 18 * np.random.default_rng(0).standard_normal(200000)), 1, D).astype(int)
probs = (np.bincount(n, minlength=D + 1) / len(n)).astype(np.float32)
epic = model == "EPiC"
def sync(): torch.cuda.synchronize(); return time.perf_counter()
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 6):
    t0 = sync()
    x0, k0, mask, nn = _abi.make_source(probs, batch, D, V, 1, it * batch, dev, discrete=not epic)
    t1 = sync()
    opts = None if epic else _abi.step_options(cfg, seed=3, first_global_jet=it * batch)
    x, k, _ = nm.generate(x0, k0, mask, ts, dt, opts)
    t2 = sync()
    kin, counts = _abi.jet_observables(x, k, mask, [1.9, 0.0, 0.0], [0.8, 0.11, 0.1], V)
    t3 = sync()
    print(f"batch {it}: source {1e3*(t1-t0):.2f} ms, generate {1e3*(t2-t1):.2f} ms, observables {1e3*(t3-t2):.2f} ms, max n {int(nn.max())}, particles {int(nn.sum())}, launches {nm.launches}")
src = synthetic.source_state(batch).to(dev)
for it in range(2):
    t1 = sync()
    x, k, _ = nm.generate(src.continuous, None if epic else src.discrete, src.mask, ts, dt, None if epic else _abi.step_options(cfg, seed=3))
    t2 = sync()
    print(f"synthetic.source_state: generate {1e3*(t2-t1):.2f} ms, max n {int(src.mask.sum(1).max())}, particles {int(src.mask.sum())}")
