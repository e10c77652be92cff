"""Repeated short generations through the C ABI must finish and reproduce the first one bit for bit.

The batch holds jets of 5, 40, 77 and 150 particles (a plain tile and a pair tile side by side).  A race in the persistent
kernel shows up here as a differing output or a failed launch; tools/tile_stress.py is the long form of the same loop and
tools/gpu_job_fresh_env.sh runs it in many fresh processes (the pair tiles once dead-locked only in the first launches of
a process: two attention hand-offs in a row on one barrier, see go_attn in csrc/kernels_tftile.cu)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["FusedParticleFormer", "ParticleFormer"])
@pytest.mark.parametrize("jets", ["all", "pair"])
def test_repeated_generations_are_bit_identical(model, jets, golden_dir):
    from mmf_b200 import _abi, synthetic
    from mmf_b200.param_spec import make_config
    from oracle import mmf_oracle as orc
    g = np.load(os.path.join(golden_dir, f"traj_{model}.npz"))
    cfg = make_config(model)
    cfg.num_timesteps = int(g["num_timesteps"])
    dev = torch.device("cuda:0")
    nm = _abi.NativeModel(cfg, synthetic.make_state_dict(cfg, flavor="wide", seed=int(g["weight_seed"])), dev)
    x0 = torch.from_numpy(g["x0"]); k0 = torch.from_numpy(g["k0"]).long(); mask = torch.from_numpy(g["mask"])
    u = synthetic.uniform_draws(cfg.num_timesteps, x0.shape[0], x0.shape[1], cfg.vocab_size, seed=int(g["u_seed"]))
    if jets == "pair":
        keep = mask.reshape(mask.shape[0], -1).sum(1) > 128
        assert int(keep.sum()) >= 1
        x0, k0, mask, u = x0[keep].contiguous(), k0[keep].contiguous(), mask[keep].contiguous(), u[:, keep].contiguous()
    x0, k0, mask, u = x0.to(dev), k0.to(dev), mask.to(dev), u.to(dev)
    ts, dt = orc.time_grid(cfg)
    opts = _abi.step_options(cfg)
    first = {}
    for rep in range(6):
        for nsteps in (1, 5, 20):
            xs, ks, _ = nm.generate(x0, k0, mask, ts[:nsteps], float(dt), opts, u=u[:nsteps])
            torch.cuda.synchronize()
            if nsteps not in first:
                first[nsteps] = (xs.clone(), ks.clone())
                assert torch.isfinite(xs).all()
            else:
                assert torch.equal(first[nsteps][0], xs) and torch.equal(first[nsteps][1], ks), (rep, nsteps)
    nm.close()
