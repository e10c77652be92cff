"""Repeat the short un-forced generations of tests/test_gpu_encoder.py many times in one process: every repeat must finish
and reproduce the first one bit for bit (a race in the tile kernel shows up as a differing output long before it crashes).
usage: python tools/tile_stress.py [repeats]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200")); sys.path.insert(0, ROOT)
import numpy as np
import torch
from mmf_b200 import _abi, synthetic
from mmf_b200.param_spec import make_config
from oracle import mmf_oracle as orc

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
bad = 0
for model in ("FusedParticleFormer", "ParticleFormer"):
    g = np.load(os.path.join(ROOT, "tests", "golden", f"traj_{model}.npz"))
    cfg = make_config(model)
    cfg.num_timesteps = int(g["num_timesteps"])
    nm = _abi.NativeModel(cfg, synthetic.make_state_dict(cfg, flavor="wide", seed=int(g["weight_seed"])), dev)
    x0 = torch.from_numpy(g["x0"]).to(dev); k0 = torch.from_numpy(g["k0"]).long().to(dev); mask = torch.from_numpy(g["mask"]).to(dev)
    sel = os.environ.get("STRESS_JETS")                  # "plain": the jets of <= 128 particles only; "pair": the others only
    if sel:
        n = mask.reshape(mask.shape[0], -1).sum(1)
        keep = (n <= 128) if sel == "plain" else (n > 128)
        x0, k0, mask = x0[keep].contiguous(), k0[keep].contiguous(), mask[keep].contiguous()
    B, D = x0.shape[:2]
    u = synthetic.uniform_draws(cfg.num_timesteps, int(g["x0"].shape[0]), D, cfg.vocab_size, seed=int(g["u_seed"]))
    if sel:
        u = u[:, keep.cpu()].contiguous()
    u = u.to(dev)
    ts, dt = orc.time_grid(cfg)
    opts = _abi.step_options(cfg)
    first = {}
    for rep in range(reps):
        for nsteps in (1, 5, 20, 100):
            t_call = time.time()
            try:
                xs, ks, _ = nm.generate(x0, k0, mask, ts[:nsteps], float(dt), opts, u=u[:nsteps])
                torch.cuda.synchronize()
            except RuntimeError as e:
                print(f"{model}: repeat {rep} nsteps {nsteps} ({time.time() - t_call:.2f} s in the call): {str(e)[:120]}", flush=True)
                sys.exit(1)
            key = nsteps
            if key not in first:
                first[key] = (xs.clone(), ks.clone())
            elif not (torch.equal(first[key][0], xs) and torch.equal(first[key][1], ks)):
                bad += 1
                d = (first[key][0] - xs).abs().max().item()
                nk = (first[key][1] != ks).sum().item()
                print(f"{model}: repeat {rep} nsteps {nsteps}: output differs from the first run (max |dx| {d:.3e}, {nk} tokens)", flush=True)
    print(f"{model}: {reps} repeats x 4 generations, {bad} differing so far", flush=True)
    nm.close()
sys.exit(1 if bad else 0)
