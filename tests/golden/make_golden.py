"""Generate the committed golden fixtures by EXECUTING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these
fixtures are the pin for ``oracle/``: inputs are seeded synthetic jets and
seeded synthetic weights (``mmf_b200.synthetic``), outputs come from the
reference classes imported behind stub modules (``ref_harness``).  Weights are
not stored (21 MB); they are re-generated from (flavor, seed) and protected by a
checksum stored in each fixture.

Fixtures written next to this file:

  encoder_<Model>_<flavor>.npz   one forward, per-jet times, ragged multiplicities
  step_cases.npz                 HybridSolver.tauleap_step on supplied (vt, logits, u), tie-free
  loss_<Model>_<mode>.npz        MultiModalFlowBridge.loss with supplied time / bridge noise / categorical uniforms
  euler_step_cases.npz           HybridSolver.euler_step (categorical jump) on supplied (vt, logits, u), tie-free
  traj_<Model>.npz               full N-step simulate_dynamics with supplied uniforms
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "multimodal-flows_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness                                           # noqa: E402
from mmf_b200 import synthetic                               # noqa: E402
from mmf_b200.param_spec import make_config                  # noqa: E402
from oracle import mmf_oracle as orc                         # noqa: E402

ENC_N = [1, 7, 33, 64, 129, 150]
TRAJ_N = [5, 40, 77, 150]


def ragged_state(ns, seed, vocab=9, D=150):
    n = torch.tensor(ns, dtype=torch.int64)
    mask = synthetic.prefix_masks(n, D)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(len(ns), D, 3, generator=g) * mask
    k = torch.randint(1, vocab, (len(ns), D, 1), generator=g) * mask
    t = torch.rand(len(ns), generator=g)
    return t, x, k, mask


def ref_model(ref, cfg, sd, bridge="mmf"):
    cls = ref.MultiModalFlowBridge if bridge == "mmf" else ref.ConditionalFlowMatching
    m = cls(cfg).eval()
    m.model.load_state_dict(sd, strict=True)      # also proves key/shape equality of param_spec
    return m


def relerr(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gen_encoders(ref):
    for model in ("ParticleFormer", "FusedParticleFormer", "EPiC"):
        for flavor in ("default", "wide"):
            cfg = make_config(model)
            sd = synthetic.make_state_dict(cfg, flavor=flavor, seed=11)
            t, x, k, mask = ragged_state(ENC_N, seed=21)
            bridge = "cfm" if model == "EPiC" else "mmf"
            m = ref_model(ref, cfg, sd, bridge)
            with torch.no_grad():
                if model == "EPiC":
                    st = ref.TensorMultiModal(time=t, continuous=x.clone(), mask=mask)
                    vt = m(st)
                    logits = torch.zeros(0)
                    ovt = orc.encoder_forward(sd, cfg, t, x, None, mask)
                    ologits = logits
                else:
                    st = ref.TensorMultiModal(time=t, continuous=x.clone(), discrete=k.clone(), mask=mask)
                    vt, logits = m(st)
                    ovt, ologits = orc.encoder_forward(sd, cfg, t, x, k, mask)
            real = mask.bool().squeeze(-1)
            # SURVEY 8(c) L1 calibration: the reference's own bf16-autocast error against its fp32 self on this fixture
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                if model == "EPiC":
                    avt = m(ref.TensorMultiModal(time=t, continuous=x.clone(), mask=mask)).float()
                    alogits = logits
                else:
                    avt, alogits = m(ref.TensorMultiModal(time=t, continuous=x.clone(), discrete=k.clone(), mask=mask))
                    avt, alogits = avt.float(), alogits.float()
            ac = {"autocast_vt_rel": relerr(avt[real], vt[real]),
                  "autocast_vt_maxabs": float((avt[real] - vt[real]).abs().max() / vt[real].abs().max())}
            if model != "EPiC":
                ac["autocast_logits_rel"] = relerr(alogits[real], logits[real])
                ac["autocast_logits_maxabs"] = float((alogits[real] - logits[real]).abs().max() / logits[real].abs().max())
            print("   bf16-autocast of the reference vs its fp32 self:", {k_: f"{v_:.2e}" for k_, v_ in ac.items()})
            print(f"encoder {model:20s} {flavor:7s} |vt|max={vt[real].abs().max():.3f} "
                  f"oracle-vs-ref vt {relerr(ovt[real], vt[real]):.2e}"
                  + ("" if model == "EPiC" else f" logits {relerr(ologits[real], logits[real]):.2e} "
                     f"logit-std={logits[real].std():.3f}"))
            np.savez_compressed(
                os.path.join(HERE, f"encoder_{model}_{flavor}.npz"),
                time=t.numpy(), continuous=x.numpy(), discrete=k.numpy().astype(np.int64),
                mask=mask.numpy().astype(np.int64), vt=vt.numpy(), logits=logits.numpy(),
                weight_seed=11, weight_checksum=synthetic.state_dict_checksum(sd), **ac)


def tie_free_uniforms(lam64, g, margin=2e-5):
    """u with |u - threshold| > margin for both thresholds of every channel."""
    u = torch.rand(lam64.shape, generator=g)
    e = torch.exp(-lam64)
    thr = torch.stack([e, e * (1.0 + lam64)])
    for _ in range(50):
        bad = ((u.double()[None] - thr).abs() < margin + 1e-4 * thr).any(0)
        if not bad.any():
            return u
        u = torch.where(bad, torch.rand(lam64.shape, generator=g), u)
    raise RuntimeError("could not draw tie-free uniforms")


def gen_steps(ref):
    install_mods = sys.modules
    from model.MJB import RandomTelegraphBridge           # type: ignore
    from utils.thermostats import ConstantThermostat      # type: ignore
    B, D, V = 8, 128, 9
    cases = [
        dict(T=1.0, top_k=None, top_p=None),
        dict(T=0.8, top_k=None, top_p=None),
        dict(T=1.2, top_k=None, top_p=None),
        dict(T=1.0, top_k=5, top_p=None),
        dict(T=0.9, top_k=None, top_p=0.9),
        dict(T=1.1, top_k=4, top_p=0.8),
    ]
    out = {}
    for ci, case in enumerate(cases):
        g = torch.Generator().manual_seed(100 + ci)
        cfg = make_config("ParticleFormer", temperature=case["T"], top_k=case["top_k"], top_p=case["top_p"])
        vt = torch.randn(B, D, 3, generator=g) * 2.0
        logits = torch.randn(B, D, V, generator=g) * 2.5
        x = torch.randn(B, D, 3, generator=g)
        k = torch.randint(0, V, (B, D, 1), generator=g)
        # times across the whole grid including the blown-up tail
        t = torch.tensor([1e-5, 0.1, 0.35, 0.6, 0.85, 0.97, 0.99, 1.0 - 1e-5], dtype=torch.float32)
        dt = torch.tensor(0.010100808, dtype=torch.float32)

        class Stub:
            bridge_discrete = RandomTelegraphBridge(cfg.beta, V, ConstantThermostat(cfg.beta, V))

            def eval(self):
                return self

            def __call__(self, state):
                return vt.clone(), logits.clone()

        # pass 1: reference rates (independent of u) to place tie-free uniforms
        solver = ref.HybridSolver(model=Stub(), config=cfg)
        st = ref.TensorMultiModal(time=t, continuous=x.clone(), discrete=k.clone(), mask=torch.ones(B, D, 1, dtype=torch.int64))
        with ref_harness.supplied_uniforms([torch.full((B, D, V), 0.5)]):
            _, rates0 = solver.tauleap_step(st, dt)
        if case["top_p"] is not None:      # keep the nucleus cut away from fp ties
            p = torch.softmax(logits.double() / case["T"], -1)
            if case["top_k"] is not None:
                p = orc.top_k_filter(p, case["top_k"], V)
            cum = torch.sort(p, -1, descending=True)[0].cumsum(-1)
            assert ((cum - case["top_p"]).abs() > 1e-6).all(), "regenerate: top-p tie"
        u = tie_free_uniforms((rates0.double() * dt.double()), g)
        st = ref.TensorMultiModal(time=t, continuous=x.clone(), discrete=k.clone(), mask=torch.ones(B, D, 1, dtype=torch.int64))
        with ref_harness.supplied_uniforms([u]):
            st2, rates = solver.tauleap_step(st, dt)
        ox, ok, orates = orc.hybrid_step(vt, logits, x, k, t, dt, u, temperature=case["T"], beta=cfg.beta,
                                         vocab_size=V, top_k=case["top_k"], top_p=case["top_p"])
        changed = float((st2.discrete != k).float().mean())
        print(f"step case {ci} {case}: oracle k equal={torch.equal(ok, st2.discrete)} x equal={torch.equal(ox, st2.continuous)} "
              f"rates rel={relerr(orates, rates):.1e} changed={changed:.3f}")
        pre = f"c{ci}_"
        out.update({pre + "vt": vt.numpy(), pre + "logits": logits.numpy(), pre + "x": x.numpy(),
                    pre + "k": k.numpy().astype(np.uint8), pre + "t": t.numpy(), pre + "dt": dt.numpy(),
                    pre + "u": u.numpy(), pre + "T": np.float32(case["T"]),
                    pre + "top_k": np.int32(case["top_k"] or 0), pre + "top_p": np.float32(case["top_p"] or 0.0),
                    pre + "x_out": st2.continuous.numpy(), pre + "k_out": st2.discrete.numpy().astype(np.uint8),
                    pre + "rates": rates.numpy()})
    out["num_cases"] = np.int32(len(cases))
    out["beta"] = np.float32(0.075)
    np.savez_compressed(os.path.join(HERE, "step_cases.npz"), **out)


def gen_euler_steps(ref):
    """HybridSolver.euler_step (reference model/solvers.py:62-91) on supplied (vt, logits) with one supplied uniform per
    particle routed through Categorical.sample; uniforms kept away from the cumulative thresholds."""
    from model.MJB import RandomTelegraphBridge           # type: ignore
    from utils.thermostats import ConstantThermostat      # type: ignore
    B, D, V = 8, 128, 9
    cases = [dict(top_k=None, top_p=None), dict(top_k=5, top_p=None), dict(top_k=None, top_p=0.9)]
    out = {}
    for ci, case in enumerate(cases):
        g = torch.Generator().manual_seed(300 + ci)
        cfg = make_config("ParticleFormer", temperature=1.0, top_k=case["top_k"], top_p=case["top_p"])
        vt = torch.randn(B, D, 3, generator=g) * 2.0
        logits = torch.randn(B, D, V, generator=g) * 2.5
        x = torch.randn(B, D, 3, generator=g)
        k = torch.randint(0, V, (B, D, 1), generator=g)
        t = torch.tensor([1e-5, 0.1, 0.35, 0.6, 0.85, 0.97, 0.99, 1.0 - 1e-5], dtype=torch.float32)
        if case["top_k"] is not None or case["top_p"] is not None:
            # the filters sort the transition probabilities; in the tail of the grid they all clamp to 1 and the order of
            # ties is implementation-defined, so the filtered cases stay where the values are distinct
            t = torch.tensor([1e-5, 0.05, 0.1, 0.2, 0.35, 0.5, 0.6, 0.7], dtype=torch.float32)
        dt = torch.tensor(0.010100808, dtype=torch.float32)

        class Stub:
            bridge_discrete = RandomTelegraphBridge(cfg.beta, V, ConstantThermostat(cfg.beta, V))

            def eval(self):
                return self

            def __call__(self, state):
                return vt.clone(), logits.clone()

        # thresholds in double from the oracle's restatement, to place the uniforms
        _, _, rates64 = orc.hybrid_euler_step(vt.double(), logits.double(), x.double(), k, t.double(), dt.double(),
                                              torch.zeros(B, D, dtype=torch.float64), beta=cfg.beta, vocab_size=V)
        dp = (rates64 * dt.double()).clamp(max=1.0).scatter(-1, k, 0.0)
        dp = dp.scatter(-1, k, (1.0 - dp.sum(-1, keepdim=True)).clamp(min=0.0))
        if case["top_k"] is not None or case["top_p"] is not None:
            # the kept set must not depend on how ties are ordered: filter the channel-reversed row as well
            def filt(q):
                if case["top_k"] is not None:
                    q = orc.top_k_filter(q, case["top_k"], V)
                if case["top_p"] is not None:
                    q = orc.top_p_filter(q, case["top_p"])
                return q
            assert torch.equal(filt(dp) > 0, filt(dp.flip(-1)).flip(-1) > 0), "regenerate: outcome depends on tie order"
        if case["top_k"] is not None:
            dp = orc.top_k_filter(dp, case["top_k"], V)
        if case["top_p"] is not None:
            srt = torch.sort(dp, -1, descending=True)[0].cumsum(-1)
            assert ((srt - case["top_p"]).abs() > 1e-6).all(), "regenerate: top-p tie"
            dp = orc.top_p_filter(dp, case["top_p"])
        cum = (dp / dp.sum(-1, keepdim=True)).cumsum(-1)
        u = torch.rand(B, D, generator=g)
        for _ in range(50):
            bad = ((u.double().unsqueeze(-1) - cum).abs() < 2e-5).any(-1)
            if not bad.any():
                break
            u = torch.where(bad, torch.rand(B, D, generator=g), u)
        assert not bad.any()
        solver = ref.HybridSolver(model=Stub(), config=cfg)
        st = ref.TensorMultiModal(time=t, continuous=x.clone(), discrete=k.clone(), mask=torch.ones(B, D, 1, dtype=torch.int64))
        with ref_harness.supplied_categorical([u]):
            st2, rates = solver.euler_step(st, dt)
        ox, ok, orates = orc.hybrid_euler_step(vt, logits, x, k, t, dt, u, beta=cfg.beta, vocab_size=V, top_k=case["top_k"], top_p=case["top_p"])
        changed = float((st2.discrete != k).float().mean())
        print(f"euler case {ci} {case}: oracle k equal={torch.equal(ok, st2.discrete)} x equal={torch.equal(ox, st2.continuous)} "
              f"rates rel={relerr(orates, rates):.1e} changed={changed:.3f}")
        pre = f"c{ci}_"
        out.update({pre + "vt": vt.numpy(), pre + "logits": logits.numpy(), pre + "x": x.numpy(), pre + "k": k.numpy().astype(np.uint8),
                    pre + "t": t.numpy(), pre + "dt": dt.numpy(), pre + "u": u.numpy(), pre + "top_k": np.int32(case["top_k"] or 0),
                    pre + "top_p": np.float32(case["top_p"] or 0.0), pre + "x_out": st2.continuous.numpy(),
                    pre + "k_out": st2.discrete.numpy().astype(np.uint8), pre + "rates": rates.numpy()})
    out["num_cases"] = np.int32(len(cases))
    out["beta"] = np.float32(0.075)
    np.savez_compressed(os.path.join(HERE, "euler_step_cases.npz"), **out)


def gen_loss(ref):
    """MultiModalFlowBridge.loss (reference model/MMF.py:138-170) with its three draws supplied: time (torch.rand),
    bridge noise z (torch.randn_like) and the categorical uniforms of RandomTelegraphBridge.sample."""
    for model, mt in (("FusedParticleFormer", "time-weighted"), ("ParticleFormer", "time-weighted"), ("FusedParticleFormer", "sum")):
        cfg = make_config(model, multitask_loss=mt, sigma=1e-3)
        sd = synthetic.make_state_dict(cfg, flavor="wide", seed=13)
        g = torch.Generator().manual_seed(400)
        ns = [1, 9, 40, 77, 128, 150]
        B, D, V, E = len(ns), 150, cfg.vocab_size, cfg.n_embd
        mask = synthetic.prefix_masks(torch.tensor(ns), D)
        x0 = torch.randn(B, D, 3, generator=g) * mask
        k0 = torch.randint(1, V, (B, D, 1), generator=g) * mask
        x1 = (torch.randn(B, D, 3, generator=g) * 1.5 + 0.3) * mask
        k1 = torch.randint(1, V, (B, D, 1), generator=g) * mask
        u01 = torch.rand(B, generator=g)
        z = torch.randn(B, D, 3, generator=g)
        sd_loss = {}
        if mt == "time-weighted":
            sd_loss = {"uncertainty_net.c_fc.weight": torch.randn(E, E, generator=g) * 0.05, "uncertainty_net.c_fc.bias": torch.randn(E, generator=g) * 0.05,
                       "uncertainty_net.c_proj.weight": torch.randn(2, E, generator=g) * 0.2, "uncertainty_net.c_proj.bias": torch.randn(2, generator=g) * 0.3}
        m = ref_model(ref, cfg, sd, "mmf")
        m.loss_combine.load_state_dict(sd_loss, strict=True)
        t = cfg.time_eps + (1.0 - cfg.time_eps) * u01
        # categorical uniforms away from the cumulative thresholds of the bridge probabilities (evaluated in double)
        kk = torch.arange(V).view(1, 1, -1).expand(B, D, -1).double()
        def cp(ti, to, ki, ko):
            w = torch.exp(-V * cfg.beta * (torch.as_tensor(to, dtype=torch.float64) - torch.as_tensor(ti, dtype=torch.float64)).expand(B))
            return 1.0 / V + w[:, None, None] * ((-1.0 / V) + (ko == ki).double())
        p = cp(t.double(), 1.0, kk, k1.double()) * cp(0.0, t.double(), k0.double(), kk) / cp(0.0, 1.0, k0.double(), k1.double())
        cum = (p / p.sum(-1, keepdim=True)).cumsum(-1)
        u = torch.rand(B, D, generator=g)
        for _ in range(50):
            bad = ((u.double().unsqueeze(-1) - cum).abs() < 2e-5).any(-1)
            if not bad.any():
                break
            u = torch.where(bad, torch.rand(B, D, generator=g), u)
        assert not bad.any()
        batch = ref.DataCoupling(source=ref.TensorMultiModal(continuous=x0.clone(), discrete=k0.clone(), mask=mask),
                                 target=ref.TensorMultiModal(continuous=x1.clone(), discrete=k1.clone(), mask=mask))
        with torch.no_grad(), ref_harness.supplied_rand(u01, z) as used, ref_harness.supplied_categorical([u], module="model.MJB"):
            out = m.loss(batch)
        assert used == {"rand": 1, "randn_like": 1}, used
        oo = orc.training_loss(sd, sd_loss, cfg, x0, k0, x1, k1, mask, t, z, u)
        vals = [float(v) if v is not None else float("nan") for v in out]
        ovals = [float(v) if v is not None else float("nan") for v in oo[:5]]
        print(f"loss {model} {mt}: reference {vals}  oracle {ovals}")
        np.savez_compressed(os.path.join(HERE, f"loss_{model}_{mt}.npz"), x0=x0.numpy(), k0=k0.numpy().astype(np.uint8), x1=x1.numpy(),
                            k1=k1.numpy().astype(np.uint8), mask=mask.numpy(), u01=u01.numpy(), time=t.numpy(), z=z.numpy(), u=u.numpy(),
                            xt=oo[5].numpy(), kt=oo[6].numpy().astype(np.uint8), out=np.array(vals, np.float32), weight_seed=13, sigma=np.float32(cfg.sigma),
                            weight_checksum=synthetic.state_dict_checksum(sd), **{"net_" + k.replace(".", "_"): v.numpy() for k, v in sd_loss.items()})


def gen_trajectories(ref):
    for model, N in (("FusedParticleFormer", 100), ("ParticleFormer", 100), ("EPiC", 100)):
        cfg = make_config(model, num_timesteps=N, temperature=1.0)
        sd = synthetic.make_state_dict(cfg, flavor="wide", seed=12)
        _, x0, k0, mask = ragged_state(TRAJ_N, seed=22)
        B, D = x0.shape[:2]
        t0 = time.time()
        if model == "EPiC":
            m = ref_model(ref, cfg, sd, "cfm")
            batch = ref.DataCoupling(source=ref.TensorMultiModal(continuous=x0.clone(), mask=mask), target=ref.TensorMultiModal())
            out = m.simulate_dynamics(batch).target
            ox = orc.simulate_dynamics_cfm(sd, cfg, x0, mask)
            real = mask.bool().squeeze(-1)
            print(f"traj {model}: ref {time.time()-t0:.1f}s  oracle-vs-ref x rel={relerr(ox[real], out.continuous[real]):.2e}")
            np.savez_compressed(os.path.join(HERE, f"traj_{model}.npz"), x0=x0.numpy(), mask=mask.numpy(),
                                x_out=out.continuous.numpy(), num_timesteps=N, weight_seed=12,
                                weight_checksum=synthetic.state_dict_checksum(sd))
            continue
        u = synthetic.uniform_draws(N, B, D, cfg.vocab_size, seed=1237)
        m = ref_model(ref, cfg, sd, "mmf")
        traj = []
        real_step = ref.HybridSolver.tauleap_step

        def spy(self, state, delta_t):
            s, r = real_step(self, state, delta_t)
            traj.append(s.discrete.clone())
            return s, r

        ref.HybridSolver.tauleap_step = spy
        try:
            batch = ref.DataCoupling(source=ref.TensorMultiModal(continuous=x0.clone(), discrete=k0.clone(), mask=mask),
                                     target=ref.TensorMultiModal())
            with ref_harness.supplied_uniforms(list(u)):
                out = m.simulate_dynamics(batch).target
        finally:
            ref.HybridSolver.tauleap_step = real_step
        traj = torch.stack(traj)
        ox, ok, _, otraj = orc.simulate_dynamics(sd, cfg, x0, k0, mask, u=u, return_trajectory=True)
        real = mask.bool().squeeze(-1)
        agree = float((otraj[:, real] == traj[:, real]).float().mean())
        print(f"traj {model}: ref {time.time()-t0:.1f}s  oracle-vs-ref x rel={relerr(ox[real], out.continuous[real]):.2e} "
              f"k equal={torch.equal(ok[real], out.discrete[real])} traj agreement={agree:.6f} "
              f"final-token-0 frac={(out.discrete[real]==0).float().mean():.3f}")
        np.savez_compressed(os.path.join(HERE, f"traj_{model}.npz"), x0=x0.numpy(), k0=k0.numpy().astype(np.uint8),
                            mask=mask.numpy(), x_out=out.continuous.numpy(), k_out=out.discrete.numpy().astype(np.uint8),
                            traj_k=traj.numpy().astype(np.uint8), num_timesteps=N, u_seed=1237,
                            u_checksum=float(u.double().sum()), weight_seed=12,
                            weight_checksum=synthetic.state_dict_checksum(sd))


if __name__ == "__main__":
    torch.manual_seed(0)
    ref = ref_harness.modules()
    which = sys.argv[1:] or ["encoders", "steps", "euler", "loss", "traj"]
    if "encoders" in which:
        gen_encoders(ref)
    if "steps" in which:
        gen_steps(ref)
    if "euler" in which:
        gen_euler_steps(ref)
    if "loss" in which:
        gen_loss(ref)
    if "traj" in which:
        gen_trajectories(ref)
