"""Parameter inventories of the three encoders on the hot path.

The drop-in promise is that a reference checkpoint loads unchanged, so the
``state_dict`` keys and shapes below must equal the reference's:

  * ParticleFormer        reference ``networks/ParticleTransformers.py:19-60``
  * FusedParticleFormer   reference ``networks/ParticleTransformers.py:146-175``
  * EPiC                  reference ``networks/EPiC.py:10-35, 96-107, 145-148``
  * attention block       reference ``networks/attention.py:6-21, 32-51``, ``utils/models.py:9-18, 31-34``

Instead of re-building the reference's module classes, each encoder is described
by a flat table ``[(dotted_name, shape, kind)]``.  ``kind`` drives initialisation
and the test-fixture generator:

  "w"   matrix weight   N(0, 0.02^2)           (reference ``_init_weights``)
  "b"   linear bias     0
  "g"   LayerNorm gain  1
  "s"   LayerNorm shift 0
  "wn_v"/"wn_g"  weight-norm direction / magnitude of EPiC linears
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Tuple

Spec = List[Tuple[str, Tuple[int, ...], str]]

DEFAULTS = dict(            # reference scripts/train_mmf.py:31-61
    vocab_size=9, dim_continuous=3, n_embd=256, n_inner=512, n_layer=5, n_layer_fused=6,
    n_head=4, dropout=0.0, qk_layernorm=True, bias=True, use_coocurrence=False,
    multitask_loss="time-weighted", beta=0.075, sigma=1e-5, time_eps=1e-5,
    num_timesteps=100, temperature=1.0, top_k=None, top_p=None, use_final_max_rates=False,
    max_num_particles=150, n_embd_glob=16,
)


def make_config(model: str = "ParticleFormer", **overrides) -> SimpleNamespace:
    """Namespace carrying every field the hot path reads."""
    cfg = dict(DEFAULTS)
    cfg["model"] = model
    cfg["metadata"] = {"mean": [0.0, 0.0, 0.0], "std": [1.0, 1.0, 1.0]}
    cfg.update(overrides)
    return SimpleNamespace(**cfg)


def _linear(spec: Spec, name: str, n_out: int, n_in: int, bias: bool = True) -> None:
    spec.append((f"{name}.weight", (n_out, n_in), "w"))
    if bias:
        spec.append((f"{name}.bias", (n_out,), "b"))


def _layernorm(spec: Spec, name: str, width: int, bias: bool = True) -> None:
    spec.append((f"{name}.weight", (width,), "g"))
    if bias:
        spec.append((f"{name}.bias", (width,), "s"))


def _attn_block(spec: Spec, name: str, cfg, width: int) -> None:
    inner = cfg.n_inner if cfg.n_inner is not None else 4 * width
    hs = width // cfg.n_head
    _layernorm(spec, f"{name}.ln1", width, cfg.bias)
    _linear(spec, f"{name}.attn.c_attn", 3 * width, width, cfg.bias)
    _linear(spec, f"{name}.attn.c_proj", width, width, cfg.bias)
    if cfg.qk_layernorm:
        _layernorm(spec, f"{name}.attn.q_layernorm", hs, cfg.bias)
        _layernorm(spec, f"{name}.attn.k_layernorm", hs, cfg.bias)
    _layernorm(spec, f"{name}.ln2", width, cfg.bias)
    _linear(spec, f"{name}.ffw.c_fc", inner, width, cfg.bias)
    _linear(spec, f"{name}.ffw.c_proj", width, inner, cfg.bias)


def _embed_and_heads(spec_head: Spec, spec_tail: Spec, cfg) -> None:
    E, h, V, dc, I = cfg.n_embd, cfg.n_embd // 2, cfg.vocab_size, cfg.dim_continuous, cfg.n_inner
    t = "transformer"
    _linear(spec_head, f"{t}.wxe.0", E, dc)
    _linear(spec_head, f"{t}.wxe.2", h, E)
    spec_head.append((f"{t}.wye.0.weight", (V, E), "w"))
    _linear(spec_head, f"{t}.wye.2", h, E)
    _layernorm(spec_head, f"{t}.ln1_x", h)
    _layernorm(spec_head, f"{t}.ln1_y", h)
    _linear(spec_tail, f"{t}.head_x.0", I, h)
    _linear(spec_tail, f"{t}.head_x.2", dc, I)
    _linear(spec_tail, f"{t}.head_y.0", I, h)
    _linear(spec_tail, f"{t}.head_y.2", V, I)


def particleformer_spec(cfg) -> Spec:
    h = cfg.n_embd // 2
    t = "transformer"
    spec: Spec = []
    tail: Spec = []
    _embed_and_heads(spec, tail, cfg)
    for i in range(cfg.n_layer):
        _attn_block(spec, f"{t}.blocks_x.{i}", cfg, h)
    for i in range(cfg.n_layer):
        _attn_block(spec, f"{t}.blocks_y.{i}", cfg, h)
    _layernorm(spec, f"{t}.ln2_x", h)
    _layernorm(spec, f"{t}.ln2_y", h)
    for i in range(cfg.n_layer_fused):
        _attn_block(spec, f"{t}.blocks_fuse.{i}", cfg, cfg.n_embd)
    _linear(spec, f"{t}.time_expand", cfg.n_embd, h)
    _layernorm(spec, f"{t}.ln3_x", h)
    _layernorm(spec, f"{t}.ln3_y", h)
    return spec + tail


def fused_particleformer_spec(cfg) -> Spec:
    t = "transformer"
    spec: Spec = []
    tail: Spec = []
    _embed_and_heads(spec, tail, cfg)
    for i in range(cfg.n_layer):
        _attn_block(spec, f"{t}.blocks.{i}", cfg, cfg.n_embd)
    _layernorm(spec, f"{t}.ln2", cfg.n_embd)
    return spec + tail


def _wn_linear(spec: Spec, name: str, n_out: int, n_in: int) -> None:
    # torch.nn.utils.weight_norm: weight = g * v / ||v||_row   (reference EPiC.py:4)
    spec.append((f"{name}.bias", (n_out,), "b"))
    spec.append((f"{name}.weight_g", (n_out, 1), "wn_g"))
    spec.append((f"{name}.weight_v", (n_out, n_in), "wn_v"))


def epic_spec(cfg) -> Spec:
    E, G, dc = cfg.n_embd, cfg.n_embd_glob, cfg.dim_continuous
    spec: Spec = []
    _linear(spec, "epic.wxe", E, dc)
    _wn_linear(spec, "epic.proj.mlp_local.0", E, 2 * E)
    _wn_linear(spec, "epic.proj.mlp_local.2", E, E)
    _wn_linear(spec, "epic.proj.mlp_global.0", E, 3 * E)
    _wn_linear(spec, "epic.proj.mlp_global.2", G, E)
    for i in range(cfg.n_layer):
        p = f"epic.layers.{i}"
        _wn_linear(spec, f"{p}.fc_glob1", E, 2 * E + G)
        _wn_linear(spec, f"{p}.fc_glob2", G, E)
        _wn_linear(spec, f"{p}.fc_loc1", E, 2 * E + G)
        _wn_linear(spec, f"{p}.fc_loc2", E, E)
    _linear(spec, "epic.head", dc, 2 * E + G)
    return spec


SPEC_BUILDERS = {
    "ParticleFormer": particleformer_spec,
    "FusedParticleFormer": fused_particleformer_spec,
    "EPiC": epic_spec,
}


def spec_for(cfg) -> Spec:
    try:
        return SPEC_BUILDERS[cfg.model](cfg)
    except KeyError:
        raise KeyError(
            f"model '{cfg.model}' is not on the accelerated path; supported: {sorted(SPEC_BUILDERS)}"
        ) from None


def count_params(spec: Spec) -> int:
    total = 0
    for _, shape, _ in spec:
        n = 1
        for s in shape:
            n *= s
        total += n
    return total
