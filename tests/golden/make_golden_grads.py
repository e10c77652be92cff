"""Gradient goldens of the training step, written by EXECUTING THE REFERENCE ITSELF (build container only):

    python tests/golden/make_golden_grads.py

For every loss fixture (tests/golden/loss_<Model>_<mode>.npz: inputs, the three supplied draws, the loss-net weights) the
reference's own ``MultiModalFlowBridge.loss(batch)[0].backward()`` (model/MMF.py:138-170 under torch autograd, fp32) is run and,
for every parameter of ``model.*`` and ``loss_combine.*``, the gradient's L2 norm and 48 entries at seeded positions are stored in
``grad_<Model>_<mode>.npz`` (the full gradients are 22 MB).  ``tests/test_oracle_golden.py`` holds autograd over the oracle
restatement to them; the GPU tests then compare the kernels with that autograd on the whole gradient.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg                                      # noqa: E402  (path setup + helpers)
import ref_harness                                           # noqa: E402
from mmf_b200 import synthetic                               # noqa: E402
from mmf_b200.param_spec import make_config                  # noqa: E402

SAMPLES = 48


def sample_positions(name: str, numel: int) -> np.ndarray:
    g = np.random.default_rng(abs(hash_name(name)) % (2 ** 32))
    return g.integers(0, numel, size=min(SAMPLES, numel))


def hash_name(name: str) -> int:
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def main():
    ref = ref_harness.modules()
    for model, mode in (("FusedParticleFormer", "time-weighted"), ("ParticleFormer", "time-weighted"), ("FusedParticleFormer", "sum")):
        g = np.load(os.path.join(HERE, f"loss_{model}_{mode}.npz"))
        cfg = make_config(model, multitask_loss=mode, sigma=float(g["sigma"]))
        sd = synthetic.make_state_dict(cfg, flavor="wide", seed=int(g["weight_seed"]))
        sd_loss = {k[4:].replace("uncertainty_net_", "uncertainty_net.").replace("c_fc_", "c_fc.").replace("c_proj_", "c_proj."): torch.from_numpy(g[k])
                   for k in g.files if k.startswith("net_")}
        T = lambda n: torch.from_numpy(g[n])
        m = mg.ref_model(ref, cfg, sd, "mmf").train()
        m.loss_combine.load_state_dict(sd_loss, strict=True)
        mask = T("mask")
        batch = ref.DataCoupling(source=ref.TensorMultiModal(continuous=T("x0").clone(), discrete=T("k0").long(), mask=mask),
                                 target=ref.TensorMultiModal(continuous=T("x1").clone(), discrete=T("k1").long(), mask=mask))
        with ref_harness.supplied_rand(T("u01"), T("z")) as used, ref_harness.supplied_categorical([T("u")], module="model.MJB"):
            out = m.loss(batch)
        assert used == {"rand": 1, "randn_like": 1}, used
        assert abs(float(out[0]) - float(g["out"][0])) <= 1e-6 * abs(float(g["out"][0]))
        out[0].backward()
        store = {"loss": np.float32(float(out[0]))}
        names = []
        for prefix, mod in (("model.", m.model), ("loss_combine.", m.loss_combine)):
            for n, p in mod.named_parameters():
                name = prefix + n
                names.append(name)
                gr = p.grad.detach().double().flatten()
                pos = sample_positions(name, gr.numel())
                key = name.replace(".", "/")
                store["norm:" + key] = np.float64(float(gr.norm()))
                store["val:" + key] = gr[torch.from_numpy(pos)].numpy().astype(np.float32)
        store["names"] = np.array(names)
        np.savez_compressed(os.path.join(HERE, f"grad_{model}_{mode}.npz"), **store)
        tot = sum(float(store["norm:" + n.replace(".", "/")]) ** 2 for n in names) ** 0.5
        print(f"grad {model} {mode}: {len(names)} parameters, |grad| = {tot:.6f}, loss = {float(out[0]):.6f}")


if __name__ == "__main__":
    main()
