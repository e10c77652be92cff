"""BASELINE.json configs #3 / #4 on the local GPU(s): FusedParticleFormer at {100,500,1000} steps x T {0.8,1.0,1.2} and
EPiC at a large batch.  Prints one JSON line per case (device-timed jets/s, CUDA events, 2 timed repeats after 1 warm-up)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from mmf_b200.mmf import time_grid

dev = torch.device("cuda:0")
def run(model, B, N, T, reps=2):
    cfg = make_config(model, num_timesteps=N, temperature=T)
    sd = synthetic.make_state_dict(cfg, "wide", 0)
    nm = _abi.NativeModel(cfg, sd, dev)
    src = synthetic.source_state(B).to(dev)
    ts, dt = time_grid(cfg)
    opts = None if model == "EPiC" else _abi.step_options(cfg, seed=3)
    k0 = None if model == "EPiC" else src.discrete
    nm.generate(src.continuous, k0, src.mask, ts, dt, opts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        x, k, _ = nm.generate(src.continuous, k0, src.mask, ts, dt, opts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = {"model": model, "jets": B, "timesteps": N, "temperature": T, "ms": round(ms, 2), "jets_per_s": round(B / ms * 1e3, 1),
           "finite": bool(torch.isfinite(x).all())}
    if k is not None:
        real = src.mask.bool().squeeze(-1)
        out["token_fractions"] = [round(float((k[real] == v).float().mean()), 4) for v in range(9)]
    print(json.dumps(out), flush=True)
    nm.close()

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fused"):
    for N in (100, 500, 1000):
        for T in (0.8, 1.0, 1.2):
            run("FusedParticleFormer", 256, N, T, reps=1 if N > 100 else 2)
if which in ("all", "epic"):
    for B in (256, 4096, 16384):
        run("EPiC", B, 100, 1.0)
if which in ("all", "pf"):
    for B in (256, 1024, 4096):
        run("ParticleFormer", B, 100, 1.0, reps=1)
