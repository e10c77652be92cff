"""TEST INFRASTRUCTURE - BASELINE config #1: the reference's tutorial toy (coloured 8 Gaussians -> 2 moons), restated.

The toy lives only in ``notebooks/Tutorial_Colored_8Gaussians_to_2Moons.ipynb`` (+ ``utils/toy_data.py``, ``utils/models.py:40-59``);
it cannot be imported (the model and bridges are notebook cells, ``toy_data`` needs torchdyn), so this file restates

  * the data            ``utils/toy_data.py:6-52`` (NGaussians), ``:73-94`` (TwoMoons; ``torchdyn.generate_moons`` -> ``sklearn.make_moons``,
                        the same two half-circles with Gaussian noise)
  * the network         notebook cell 4: ``wt`` TimeFourierEmbedding, ``wx`` Linear(2,E), ``wk`` Embedding(S+1,E), 3-layer GELU MLP,
                        ``reg_head`` Linear(E/2,2), ``class_head`` Linear(E/2,S)
  * the loss            notebook cell 4 ``multimodal_loss`` with the bridges of cell 3 (tokens 1..S)
  * the sampler         notebook cell 4 ``simulate_dynamics``: grid linspace(0,1,N) (so the LAST point has w = 1, coefficient
                        w S/(1-w) = inf), softmax, telegraph rate over S classes, tau-leap with at most one jump, Euler step

``torch.poisson`` is replaced by its uniform-driven equivalent (oracle.mmf_oracle.poisson_counts_from_uniform), which makes the
sampler a function of supplied draws.  The hot-path claim tested with it (tests/test_toy_config1.py): the sampler's step is the
library's fused hybrid step at vocab_size 8, beta 0.25 - bit for bit on every grid point of a 100-step run.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .mmf_oracle import poisson_counts_from_uniform


def eight_gaussians(n_per: int, gen: torch.Generator, std_dev: float = 0.1, scale: float = 5.0):
    """utils/toy_data.py:23-51: points around 8 centres on a circle of radius `scale`, label = centre index + 1."""
    xs, ks = [], []
    for i in range(8):
        ang = i * 2 * math.pi / 8
        pts = torch.randn(n_per, 2, generator=gen) * math.sqrt(math.sqrt(std_dev))    # MultivariateNormal(0, sqrt(std) I): std = sqrt(sqrt(std_dev))
        xs.append(pts + torch.tensor([math.cos(ang), math.sin(ang)]) * scale)
        ks.append(torch.full((n_per,), i + 1))
    x, k = torch.cat(xs), torch.cat(ks)
    idx = torch.randperm(len(k), generator=gen)
    return x[idx].float(), k[idx].long()


def two_moons(n_per: int, seed: int, std_dev: float = 0.2):
    """utils/toy_data.py:88-94: moons * 3 - 1, label = moon index + 1."""
    from sklearn.datasets import make_moons
    pos, lab = make_moons(2 * n_per, noise=std_dev, random_state=seed)
    return torch.from_numpy(pos).float() * 3 - 1, torch.from_numpy(lab).long() + 1


def init_params(n_embd: int, vocab_size: int, gen: torch.Generator):
    """The notebook's module tree with torch's default initialisers drawn from `gen`."""
    E, S = n_embd, vocab_size

    def lin(o, i):
        b = 1.0 / math.sqrt(i)
        return (torch.rand(o, i, generator=gen) * 2 - 1) * b, (torch.rand(o, generator=gen) * 2 - 1) * b
    p = {}
    p["wx.w"], p["wx.b"] = lin(E, 2)
    p["wk"] = torch.randn(S + 1, E, generator=gen)
    p["wk"][0] = 0.0                                        # padding_idx = 0
    for i in (0, 2, 4):
        p[f"mlp.{i}.w"], p[f"mlp.{i}.b"] = lin(E, E)
    p["reg.w"], p["reg.b"] = lin(2, E // 2)
    p["cls.w"], p["cls.b"] = lin(S, E // 2)
    return {k: v.clone().requires_grad_(True) for k, v in p.items()}


def forward(p, t, x, k, n_embd: int):
    """notebook cell 4 `forward`: (ut (B,2), ht (B,S))."""
    half = n_embd // 2
    inv_freq = 1.0 / (10.0 ** (torch.arange(half).float() / (half - 1)))       # utils/models.py:46-50
    a = t[:, None] * inv_freq[None, :]
    t_emb = torch.cat([a.sin(), a.cos()], dim=-1)
    h = F.linear(x, p["wx.w"], p["wx.b"]) + p["wk"][k] + t_emb
    h = F.linear(F.gelu(F.linear(F.gelu(F.linear(h, p["mlp.0.w"], p["mlp.0.b"])), p["mlp.2.w"], p["mlp.2.b"])), p["mlp.4.w"], p["mlp.4.b"])
    h1, h2 = h.split((half, half), dim=-1)
    return F.linear(h1, p["reg.w"], p["reg.b"]), F.linear(h2, p["cls.w"], p["cls.b"])


def _cond_prob(t_in, t_out, k_in, k_out, beta, S):
    w = torch.exp(-S * beta * (t_out - t_in))
    delta = (k_out == k_in).float()
    return 1.0 / S + (w if delta.dim() == 1 else w[:, None]) * (delta - 1.0 / S)


def loss(p, x0, k0, x1, k1, n_embd, S, sigma, beta, gen):
    """notebook cell 4 `multimodal_loss` (draws from `gen`)."""
    B = len(x0)
    t = torch.rand(B, generator=gen)
    xt = t[:, None] * x1 + (1.0 - t[:, None]) * x0 + sigma * torch.randn(B, 2, generator=gen)
    kk = torch.arange(1, S + 1).view(1, S).expand(B, S)
    pr = (_cond_prob(t, torch.ones_like(t), kk, k1.view(-1, 1).expand_as(kk), beta, S)
          * _cond_prob(torch.zeros_like(t), t, k0.view(-1, 1).expand_as(kk), kk, beta, S)
          / _cond_prob(torch.zeros_like(t), torch.ones_like(t), k0, k1, beta, S).view(-1, 1))
    kt = torch.multinomial(pr, 1, generator=gen).view(-1) + 1
    ut, ht = forward(p, t, xt, kt, n_embd)
    return F.mse_loss(ut, x1 - x0) + F.cross_entropy(ht, (k1 - 1).clamp(min=0))


def train(p, x0, k0, x1, k1, n_embd, S, sigma, beta, steps, gen, lr=1e-3, batch=256):
    opt = torch.optim.Adam(list(p.values()), lr=lr)
    for _ in range(steps):
        i0 = torch.randint(0, len(x0), (batch,), generator=gen)
        i1 = torch.randint(0, len(x1), (batch,), generator=gen)
        opt.zero_grad()
        l = loss(p, x0[i0], k0[i0], x1[i1], k1[i1], n_embd, S, sigma, beta, gen)
        l.backward()
        opt.step()
    return float(l)


def sampler_step(ut, ht, x, k, t, dt, u, beta, S):
    """One grid point of notebook cell 4 `simulate_dynamics` after the forward: tokens 1..S, x (B,2), u (B,S)."""
    probs = F.softmax(ht, dim=-1)
    idx = (k - 1).clamp(min=0, max=S - 1)
    qy = torch.gather(probs, 1, idx.view(-1, 1))
    w = torch.exp(-S * beta * (torch.ones_like(t) - t))
    rates = 1.0 + ((w * S) / (1.0 - w))[:, None] * probs + w[:, None] * qy
    k_idx = k - 1
    dn = poisson_counts_from_uniform(rates * dt, u)
    mask = (dn.sum(-1) <= 1).to(k_idx.dtype)
    diff = torch.arange(S).view(1, S) - k_idx.unsqueeze(-1)
    net = (dn * diff).sum(-1).to(k_idx.dtype)
    k_idx = (k_idx + net * mask) % S
    return x + ut * dt, k_idx + 1, rates
