"""CPU oracle for the generation hot path of dfaroughy/Multimodal-flows.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
the timed CPU port, never as a fallback.

It is a from-scratch fp32 restatement, in plain functional PyTorch over a
``state_dict``, of exactly the functions SURVEY.md section 8(a) lists.  Every
function cites the reference file:line it follows.  Parity pinning: the
reference ships no tests or golden vectors, so this file
is pinned by *executing the reference itself* in the build container
(``tests/golden/make_golden.py``, which imports ``/root/reference`` behind stub
modules) and committing the resulting input/output vectors under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.

The only randomness on the path is ``torch.poisson`` (reference
``model/solvers.py:48``).  As in SURVEY.md section 8(a-5) it is replaced by its
uniform-driven equivalent: with one u ~ U[0,1) per destination channel,
``count = [u >= exp(-lam)] + [u >= exp(-lam) (1 + lam)]`` which is the exact
Poisson CDF inversion truncated at 2 -- all the update rule can distinguish.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------
# small pieces                                   reference utils/models.py
# --------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int, max_positions: int = 10000) -> torch.Tensor:
    """sin/cos features of the raw time in [0,1] (reference ``utils/models.py:62-75``)."""
    assert t.dim() == 1
    half = dim // 2
    scale = math.log(max_positions) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -scale)
    arg = t.float()[:, None] * freqs[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def _ln(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """LayerNorm eps=1e-5, optional shift (reference ``utils/models.py:28-37``)."""
    w = sd[f"{name}.weight"]
    return F.layer_norm(x, w.shape, w, sd.get(f"{name}.bias"), 1e-5)


def _lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[f"{name}.weight"], sd.get(f"{name}.bias"))


def _mlp(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """c_proj(GELU_erf(c_fc(x))) (reference ``utils/models.py:8-25``)."""
    return _lin(sd, f"{name}.c_proj", F.gelu(_lin(sd, f"{name}.c_fc", x)))


def _pair_mask(mask: torch.Tensor, n_head: int) -> torch.Tensor:
    """(B,D,1) int -> (B,H,D,D) bool, True = attend
    (reference ``networks/ParticleTransformers.py:64-68``; ``squeeze(-1)`` instead of the
    reference's bare ``squeeze()`` so B=1 works)."""
    m = mask.bool().squeeze(-1)
    pair = m[:, None, None, :] & m[:, None, :, None]
    return pair.expand(-1, n_head, -1, -1)


def _attention(sd: SD, name: str, x: torch.Tensor, pair: torch.Tensor, n_head: int) -> torch.Tensor:
    """Fused-QKV multi-head attention with per-head q/k LayerNorm
    (reference ``networks/attention.py:53-74``)."""
    B, T, C = x.shape
    hs = C // n_head
    q, k, v = _lin(sd, f"{name}.c_attn", x).split(C, dim=2)
    q = q.view(B, T, n_head, hs).transpose(1, 2)
    k = k.view(B, T, n_head, hs).transpose(1, 2)
    v = v.view(B, T, n_head, hs).transpose(1, 2)
    if f"{name}.q_layernorm.weight" in sd:
        q = _ln(sd, f"{name}.q_layernorm", q)
        k = _ln(sd, f"{name}.k_layernorm", k)
    y = F.scaled_dot_product_attention(q, k, v, attn_mask=pair, dropout_p=0.0, is_causal=False)
    y = y.transpose(1, 2).contiguous().view(B, T, C)
    return _lin(sd, f"{name}.c_proj", y)


def _block(sd: SD, name: str, x: torch.Tensor, pair: torch.Tensor, n_head: int) -> torch.Tensor:
    """pre-LN attention + pre-LN MLP residual block (reference ``networks/attention.py:23-26``)."""
    x = x + _attention(sd, f"{name}.attn", _ln(sd, f"{name}.ln1", x), pair, n_head)
    return x + _mlp(sd, f"{name}.ffw", _ln(sd, f"{name}.ln2", x))


def _embed_x(sd: SD, xc: torch.Tensor) -> torch.Tensor:
    t = "transformer"
    return _ln(sd, f"{t}.ln1_x", _lin(sd, f"{t}.wxe.2", F.gelu(_lin(sd, f"{t}.wxe.0", xc))))


def _embed_y(sd: SD, k: torch.Tensor) -> torch.Tensor:
    t = "transformer"
    e = F.embedding(k, sd[f"{t}.wye.0.weight"])
    return _ln(sd, f"{t}.ln1_y", _lin(sd, f"{t}.wye.2", F.gelu(e)))


def _heads(sd: SD, x: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    t = "transformer"
    vt = _lin(sd, f"{t}.head_x.2", F.gelu(_lin(sd, f"{t}.head_x.0", x)))
    logits = _lin(sd, f"{t}.head_y.2", F.gelu(_lin(sd, f"{t}.head_y.0", y)))
    return vt, logits


# --------------------------------------------------------------------------
# encoders
# --------------------------------------------------------------------------
def particleformer_forward(sd: SD, cfg, time, continuous, discrete, mask):
    """reference ``networks/ParticleTransformers.py:62-122`` (use_coocurrence=False)."""
    t = "transformer"
    h = cfg.n_embd // 2
    pair = _pair_mask(mask, cfg.n_head)
    temb = timestep_embedding(time, h).unsqueeze(1)                     # (B,1,h)

    x = _embed_x(sd, continuous) + temb
    x_skip = x
    for i in range(cfg.n_layer):
        x = _block(sd, f"{t}.blocks_x.{i}", x, pair, cfg.n_head) + temb
    x = _ln(sd, f"{t}.ln2_x", x + x_skip)

    y = _embed_y(sd, discrete.squeeze(-1)) + temb
    y_skip = y
    for i in range(cfg.n_layer):
        y = _block(sd, f"{t}.blocks_y.{i}", y, pair, cfg.n_head) + temb
    y = _ln(sd, f"{t}.ln2_y", y + y_skip)

    temb2 = _lin(sd, f"{t}.time_expand", temb)
    z = torch.cat((x, y), dim=-1) + temb2
    for i in range(cfg.n_layer_fused):
        z = _block(sd, f"{t}.blocks_fuse.{i}", z, pair, cfg.n_head) + temb2

    x, y = z.split((h, h), dim=-1)
    x = _ln(sd, f"{t}.ln3_x", x + x_skip)
    y = _ln(sd, f"{t}.ln3_y", y + y_skip)
    return _heads(sd, x, y)


def fused_particleformer_forward(sd: SD, cfg, time, continuous, discrete, mask):
    """reference ``networks/ParticleTransformers.py:177-210``."""
    t = "transformer"
    h = cfg.n_embd // 2
    pair = _pair_mask(mask, cfg.n_head)
    z = torch.cat((_embed_x(sd, continuous), _embed_y(sd, discrete.squeeze(-1))), dim=-1)
    temb = timestep_embedding(time, cfg.n_embd).unsqueeze(1)
    z = z + temb
    z_skip = z
    for i in range(cfg.n_layer):
        z = _block(sd, f"{t}.blocks.{i}", z, pair, cfg.n_head) + temb
    z = _ln(sd, f"{t}.ln2", z + z_skip)
    x, y = z.split((h, h), dim=-1)
    return _heads(sd, x, y)


def fold_weight_norm(sd: SD, name: str) -> torch.Tensor:
    """W = g * v / ||v||_row, old-style ``torch.nn.utils.weight_norm`` with dim=0
    (reference ``networks/EPiC.py:4,97-106,145-148``)."""
    v = sd[f"{name}.weight_v"]
    g = sd[f"{name}.weight_g"]
    return g * v / v.norm(dim=1, keepdim=True)


def _wn_lin(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, fold_weight_norm(sd, name), sd[f"{name}.bias"])


def _meansum_pool(mask: torch.Tensor, local: torch.Tensor, *glob: torch.Tensor, scale: float = 0.01):
    """masked mean ++ scale*sum ++ globals (reference ``networks/EPiC.py:65-72``)."""
    s = (local * mask).sum(1)
    return torch.cat([s / mask.sum(1), s * scale, *glob], dim=1)


def epic_forward(sd: SD, cfg, time, continuous, mask):
    """reference ``networks/EPiC.py:38-62`` (+ ``:110-124`` projection, ``:152-173`` layer)."""
    D = continuous.shape[1]
    maskf = mask.to(continuous.dtype)
    xe = _lin(sd, "epic.wxe", continuous)
    tglob = timestep_embedding(time, cfg.n_embd)
    tloc = tglob.unsqueeze(1).expand(-1, D, -1)

    p = "epic.proj"
    loc = F.gelu(_wn_lin(sd, f"{p}.mlp_local.2",
                         F.gelu(_wn_lin(sd, f"{p}.mlp_local.0", torch.cat([tloc, xe], dim=-1)))))
    glob = F.gelu(_wn_lin(sd, f"{p}.mlp_global.2",
                          F.gelu(_wn_lin(sd, f"{p}.mlp_global.0", _meansum_pool(maskf, loc, tglob)))))
    loc_skip, glob_skip = loc, glob

    for i in range(cfg.n_layer):
        q = f"epic.layers.{i}"
        # the layer keeps pre-activation accumulators and returns their leaky-relu
        gpre = glob + _wn_lin(sd, f"{q}.fc_glob2",
                              F.leaky_relu(_wn_lin(sd, f"{q}.fc_glob1", _meansum_pool(maskf, loc, glob))))
        g2l = gpre.unsqueeze(1).expand(-1, D, -1)
        lpre = loc + _wn_lin(sd, f"{q}.fc_loc2",
                             F.leaky_relu(_wn_lin(sd, f"{q}.fc_loc1", torch.cat([tloc, loc, g2l], dim=-1))))
        loc = F.leaky_relu(lpre) + loc_skip
        glob = F.leaky_relu(gpre) + glob_skip

    hcat = torch.cat([tloc, loc, glob.unsqueeze(1).expand(-1, D, -1)], dim=-1)
    return _lin(sd, "epic.head", hcat)


def encoder_forward(sd: SD, cfg, time, continuous, discrete, mask):
    if cfg.model == "ParticleFormer":
        return particleformer_forward(sd, cfg, time, continuous, discrete, mask)
    if cfg.model == "FusedParticleFormer":
        return fused_particleformer_forward(sd, cfg, time, continuous, discrete, mask)
    if cfg.model == "EPiC":
        return epic_forward(sd, cfg, time, continuous, mask)
    raise KeyError(cfg.model)


# --------------------------------------------------------------------------
# the hybrid step
# --------------------------------------------------------------------------
def thermostat_w(t: torch.Tensor, beta: float, vocab_size: int) -> torch.Tensor:
    """w = exp(-V beta (1 - t)) (reference ``utils/thermostats.py:20-27`` with t1 = 1)."""
    return torch.exp(-vocab_size * beta * (1.0 - t))


def telegraph_rate(t: torch.Tensor, k: torch.Tensor, probs: torch.Tensor, beta: float,
                   vocab_size: int) -> torch.Tensor:
    """rate_v = 1 + w V/(1-w) q_v + w q_k (reference ``model/MJB.py:163-195``)."""
    qk = torch.gather(probs, 2, k.long())
    w = thermostat_w(t, beta, vocab_size)
    coeff = (w * vocab_size) / (1.0 - w)
    return 1.0 + coeff[:, None, None] * probs + w[:, None, None] * qk


def top_k_filter(probs: torch.Tensor, top_k: int, vocab_size: int) -> torch.Tensor:
    """reference ``model/solvers.py:101-109``."""
    if top_k == vocab_size:
        return probs
    _, idx = torch.topk(probs, top_k, dim=-1)
    keep = torch.zeros_like(probs).scatter_(-1, idx, 1.0)
    probs = probs * keep
    return probs / (probs.sum(dim=-1, keepdim=True) + 1e-8)


def top_p_filter(probs: torch.Tensor, top_p: float) -> torch.Tensor:
    """reference ``model/solvers.py:111-119``."""
    srt, idx = torch.sort(probs, dim=-1, descending=True)
    keep = srt.cumsum(dim=-1) <= top_p
    keep[..., 0] = True
    keep = torch.zeros_like(probs).scatter(-1, idx, keep.to(probs.dtype))
    probs = probs * keep
    return probs / (probs.sum(dim=-1, keepdim=True) + 1e-8)


def poisson_counts_from_uniform(lam: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """Poisson(lam) truncated to {0,1,>=2}, by CDF inversion of one uniform."""
    e = torch.exp(-lam)
    return (u >= e).to(lam.dtype) + (u >= e * (1.0 + lam)).to(lam.dtype)


def hybrid_step(vt, logits, x, k, t, dt, u, *, temperature=1.0, beta=0.075, vocab_size=9,
                top_k=None, top_p=None):
    """One tau-leap + Euler update given the encoder outputs
    (reference ``model/solvers.py:22-60``).  ``k`` is (B,D,1) int64; returns (x', k', rates)."""
    if temperature != 1.0:
        logits = logits / temperature
    probs = F.softmax(logits, dim=-1)
    if top_k is not None:
        probs = top_k_filter(probs, top_k, vocab_size)
    if top_p is not None:
        probs = top_p_filter(probs, top_p)
    rates = telegraph_rate(t, k, probs, beta, vocab_size)
    ks = k.squeeze(-1)
    dn = poisson_counts_from_uniform(rates * dt, u)
    allow = (dn.sum(dim=-1).type_as(ks) <= 1)
    diff = torch.arange(vocab_size, device=ks.device).view(1, 1, vocab_size) - ks[:, :, None]
    net = (dn * diff).sum(dim=-1).type_as(ks)
    k_new = ((ks + net * allow) % vocab_size).unsqueeze(-1)
    x_new = x + vt * dt
    return x_new, k_new, rates


def categorical_from_uniform(probs: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """``Categorical(probs).sample()`` by inverse CDF of one supplied uniform per row (channel order):
    the index of the first channel whose cumulative normalised probability exceeds u."""
    p = probs / probs.sum(dim=-1, keepdim=True)
    cum = p.cumsum(dim=-1)
    idx = (u.unsqueeze(-1) >= cum).sum(dim=-1)
    last = (probs > 0).float().cumsum(-1).argmax(-1)            # last channel with positive mass
    return torch.minimum(idx, last)


def hybrid_euler_step(vt, logits, x, k, t, dt, u, *, beta=0.075, vocab_size=9, top_k=None, top_p=None):
    """``HybridSolver.euler_step`` for temperature 1 (reference ``model/solvers.py:62-91``): categorical jump with
    off-diagonal probabilities ``clamp(rate dt, max=1)`` and diagonal ``clamp(1 - sum, min=0)``; the filters act on
    these transition probabilities.  ``u`` is (B,D).  Returns (x', k' (B,D,1), rates)."""
    probs = F.softmax(logits, dim=-1)
    rates = telegraph_rate(t, k, probs, beta, vocab_size)
    delta_p = (rates * dt).clamp(max=1.0)
    delta_p = delta_p.scatter(-1, k, 0.0)
    delta_p = delta_p.scatter(-1, k, (1.0 - delta_p.sum(dim=-1, keepdim=True)).clamp(min=0.0))
    if top_k is not None:
        delta_p = top_k_filter(delta_p, top_k, vocab_size)
    if top_p is not None:
        delta_p = top_p_filter(delta_p, top_p)
    k_new = categorical_from_uniform(delta_p, u).unsqueeze(-1)
    return x + vt * dt, k_new, rates


# --------------------------------------------------------------------------
# forward half of the training step                 reference model/MMF.py:138-170
# --------------------------------------------------------------------------
def telegraph_conditional_probability(t_in, t_out, k_in, k_out, beta: float, vocab_size: int) -> torch.Tensor:
    """p(k_in -> k_out; t_in, t_out) = 1/V + w (delta - 1/V), w = exp(-V beta (t_out - t_in))
    (reference ``model/MJB.py:231-253`` with the constant thermostat); t_* are (B,) tensors or floats, k_* (B,D,*)."""
    t_in = torch.as_tensor(t_in, dtype=torch.float32, device=k_out.device)
    t_out = torch.as_tensor(t_out, dtype=torch.float32, device=k_out.device)
    B = k_out.shape[0]
    t_in = t_in.expand(B) if t_in.dim() == 0 else t_in
    t_out = t_out.expand(B) if t_out.dim() == 0 else t_out
    w = torch.exp(-vocab_size * beta * (t_out - t_in))
    kron = (k_out == k_in).float()
    return 1.0 / vocab_size + w[:, None, None] * ((-1.0 / vocab_size) + kron)


def bridge_sample(x0, x1, k0, k1, t, z, u, *, sigma: float, beta: float, vocab_size: int):
    """xt = t x1 + (1 - t) x0 + sigma z  (reference ``model/CFM.py:171-184``) and kt ~ Categorical(P) with
    P(k) = p(k -> k1; t, 1) p(k0 -> k; 0, t) / p(k0 -> k1; 0, 1)  (``model/MJB.py:197-229``), the categorical draw by inverse
    CDF of the supplied u (B,D).  Returns xt (B,D,3), kt (B,D,1)."""
    tt = t[:, None, None]
    xt = tt * x1 + (1.0 - tt) * x0
    xt = xt + sigma * z
    k = torch.arange(vocab_size, device=k0.device).view(1, 1, -1).expand(k0.shape[0], k0.shape[1], -1).float()
    p = (telegraph_conditional_probability(t, 1.0, k, k1, beta, vocab_size)
         * telegraph_conditional_probability(0.0, t, k0, k, beta, vocab_size)
         / telegraph_conditional_probability(0.0, 1.0, k0, k1, beta, vocab_size))
    return xt, categorical_from_uniform(p, u).unsqueeze(-1)


def multitask_loss(sd_loss: SD, cfg, vt, logits, x0, x1, k1, mask, t):
    """Masked MSE + CE per jet and their combination (reference ``model/MMF.py:152-168``, ``MultiTaskLoss`` ``:203-233``).
    ``sd_loss``: the ``loss_combine.*`` entries of the checkpoint without the prefix.  Returns (loss, mse, ce, w_mse, w_ce)."""
    B, V = x0.shape[0], cfg.vocab_size
    m = mask.to(vt.dtype)
    mse = (F.mse_loss(vt, x1 - x0, reduction="none") * m).sum(dim=[1, 2]) / m.sum(dim=[1, 2]).clamp_min(1.0)
    ce = F.cross_entropy(logits.reshape(-1, V), k1.reshape(-1), ignore_index=0, reduction="none").view(B, -1) * m.squeeze(-1)
    ce = ce.sum(dim=1) / m.squeeze(-1).sum(dim=1).clamp_min(1.0)
    if cfg.multitask_loss == "sum":
        return (mse + ce).mean(), mse.mean(), ce.mean(), None, None
    if cfg.multitask_loss == "weighted":                      # reference model/MMF.py:219-223: two learned log-variances
        u1, u2 = sd_loss["loss_weights"].unbind(-1)
        w1, w2 = torch.exp(-u1), torch.exp(-u2)
        loss = 0.5 * (u1 + w1 * mse) + 0.5 * (u2 + w2 * ce)
        return loss.mean(), mse.mean(), ce.mean(), w1.mean(), w2.mean()
    temb = timestep_embedding(t, cfg.n_embd)
    h = F.gelu(F.linear(temb, sd_loss["uncertainty_net.c_fc.weight"], sd_loss["uncertainty_net.c_fc.bias"]))
    u1, u2 = F.linear(h, sd_loss["uncertainty_net.c_proj.weight"], sd_loss["uncertainty_net.c_proj.bias"]).unbind(-1)
    w1, w2 = torch.exp(-u1), torch.exp(-u2)
    loss = 0.5 * (u1 + w1 * mse) + 0.5 * (u2 + w2 * ce)
    return loss.mean(), mse.mean(), ce.mean(), w1.mean(), w2.mean()


def training_loss(sd: SD, sd_loss: SD, cfg, x0, k0, x1, k1, mask, t, z, u):
    """``MultiModalFlowBridge.loss`` with its three random draws supplied (time, bridge noise, categorical uniforms)."""
    xt, kt = bridge_sample(x0, x1, k0, k1, t, z, u, sigma=cfg.sigma, beta=cfg.beta, vocab_size=cfg.vocab_size)
    vt, logits = encoder_forward(sd, cfg, t, xt, kt, mask)
    return multitask_loss(sd_loss, cfg, vt, logits, x0, x1, k1, mask, t) + (xt, kt)


def step_uniforms(seed: int, first_global_jet: int, num_steps: int, B: int, D: int, V: int) -> torch.Tensor:
    """The in-kernel draws of the sampler, restated: u[step, b, d, v] of the library's counter-based generator
    (``philox_uniforms`` in csrc/mmf_common.cuh) - Philox4x32-10, key = seed, counter = (global slot lo, hi, step, v // 4),
    word v % 4, u = (word >> 8) 2^-24 with global slot = (first_global_jet + b) D + d.  Feeding these to
    ``simulate_dynamics(u=...)`` reproduces a production-mode (no supplied uniforms) run of ``mmf_generate`` draw for draw."""
    import numpy as np
    from .source_oracle import philox4x32_10
    slot = (np.arange(B * D, dtype=np.uint64) + np.uint64(first_global_jet) * np.uint64(D)).reshape(1, B * D)
    lo, hi = slot & np.uint64(0xFFFFFFFF), slot >> np.uint64(32)
    step = np.arange(num_steps, dtype=np.uint64).reshape(num_steps, 1)
    out = np.empty((num_steps, B * D, V), np.float32)
    for blk in range((V + 3) // 4):
        words = philox4x32_10(np.broadcast_to(lo, (num_steps, B * D)), np.broadcast_to(hi, (num_steps, B * D)),
                              np.broadcast_to(step, (num_steps, B * D)), blk, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        for j in range(4):
            if blk * 4 + j < V:
                out[:, :, blk * 4 + j] = (words[j] >> np.uint32(8)).astype(np.float32) * np.float32(5.96046448e-08)
    return torch.from_numpy(out.reshape(num_steps, B, D, V))


def time_grid(cfg, device="cpu"):
    """t_i = linspace(eps, 1-eps, N); dt = (t_{N-1} - t_0)/(N-1) (reference ``model/MMF.py:181-184``)."""
    ts = torch.linspace(cfg.time_eps, 1.0 - cfg.time_eps, cfg.num_timesteps, device=device)
    dt = (ts[-1] - ts[0]) / (len(ts) - 1)
    return ts, dt


@torch.no_grad()
def simulate_dynamics(sd: SD, cfg, source_x, source_k, mask, u: Optional[torch.Tensor] = None,
                      generator: Optional[torch.Generator] = None, forced_k: Optional[torch.Tensor] = None,
                      return_trajectory: bool = False, max_steps: Optional[int] = None):
    """The N-step sampler for the two transformers (reference ``model/MMF.py:172-200``).

    ``u``        (N,B,D,V) supplied uniform draws (drawn from ``generator`` when None)
    ``forced_k`` (N,B,D,1) teacher-forced token trajectory: when given, the token state entering
                 step i+1 is forced_k[i] instead of the state produced here (used to compare
                 continuous trajectories across implementations whose jumps may diverge).
    Returns (x, k, rates_last[, traj_k]) ; every grid point is stepped (N*dt = 1.0101).
    """
    ts, dt = time_grid(cfg, source_x.device)
    x, k = source_x.clone(), source_k.clone()
    B, D = x.shape[:2]
    traj = []
    rates = None
    steps = len(ts) if max_steps is None else min(max_steps, len(ts))
    for i in range(steps):
        t = torch.full((B,), ts[i].item(), device=x.device)
        vt, logits = encoder_forward(sd, cfg, t, x, k, mask)
        ui = u[i] if u is not None else torch.rand(B, D, cfg.vocab_size, generator=generator, device=x.device)
        x, k, rates = hybrid_step(vt, logits, x, k, t, dt, ui, temperature=cfg.temperature, beta=cfg.beta,
                                  vocab_size=cfg.vocab_size, top_k=cfg.top_k, top_p=cfg.top_p)
        if return_trajectory:
            traj.append(k.clone())
        if forced_k is not None:
            k = forced_k[i].clone()
    if cfg.use_final_max_rates:
        k = torch.max(rates, dim=2)[1].unsqueeze(-1)
    if return_trajectory:
        return x, k, rates, torch.stack(traj)
    return x, k, rates


@torch.no_grad()
def simulate_dynamics_cfm(sd: SD, cfg, source_x, mask, max_steps: Optional[int] = None):
    """EPiC carrier: Euler ODE only (reference ``model/CFM.py:133-154``, ``model/solvers.py:139-143``)."""
    ts, dt = time_grid(cfg, source_x.device)
    x = source_x.clone()
    B = x.shape[0]
    steps = len(ts) if max_steps is None else min(max_steps, len(ts))
    for i in range(steps):
        t = torch.full((B,), ts[i].item(), device=x.device)
        x = x + encoder_forward(sd, cfg, t, x, None, mask) * dt
    return x


# --------------------------------------------------------------------------
# jet observables for the histogram-level parity check
# --------------------------------------------------------------------------
def jet_observables(x: torch.Tensor, k: torch.Tensor, mask: torch.Tensor) -> Dict[str, torch.Tensor]:
    """mass, multiplicity and token fractions over real particles.

    Kinematics follow reference ``utils/aoj.py:340-346,452-462`` (px = pT cos phi, py = pT sin phi,
    pz = pT sinh eta, E = pT cosh eta on the de-standardised features); multiplicity ``aoj.py:337``;
    token fractions as in ``utils/metrics.py:10-33``.
    """
    m = mask.squeeze(-1).to(x.dtype)
    pt, eta, phi = x[..., 0], x[..., 1], x[..., 2]
    px, py = pt * torch.cos(phi) * m, pt * torch.sin(phi) * m
    pz, e = pt * torch.sinh(eta) * m, pt * torch.cosh(eta) * m
    m2 = e.sum(1) ** 2 - px.sum(1) ** 2 - py.sum(1) ** 2 - pz.sum(1) ** 2
    ks = (k.squeeze(-1) * mask.squeeze(-1)).long()
    V = int(ks.max().item()) + 1 if ks.numel() else 1
    counts = torch.stack([((ks == v) & (mask.squeeze(-1) > 0)).sum(1) for v in range(max(V, 9))], dim=1)
    return {
        "mass": torch.sqrt(torch.clamp(m2, min=0.0)),
        "multiplicity": m.sum(1),
        "token_counts": counts,
    }
