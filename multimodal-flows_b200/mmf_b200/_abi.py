"""ctypes binding of ``libmmf_b200.so`` (C ABI declared in ``include/mmf_b200.h``).

This is the only place Python touches the native library.  PyTorch is used for
device memory and streams; tensors cross the boundary as raw pointers.  There is
no fallback: if the library is missing or a call fails, a ``RuntimeError`` is
raised with ``mmf_last_error()``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_uint32, c_uint64, c_void_p
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmf_b200.so")

ARCH = {"ParticleFormer": 0, "FusedParticleFormer": 1, "EPiC": 2}


class MmfModelDesc(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "arch", "vocab_size", "dim_continuous", "n_embd", "n_inner", "n_head", "n_layer", "n_layer_fused",
        "n_embd_glob", "qk_layernorm", "max_num_particles")]


class MmfWeightRef(ctypes.Structure):
    _fields_ = [("name", c_char_p), ("data", c_void_p), ("ndim", c_int32), ("shape", c_int64 * 4)]


class MmfStepOptions(ctypes.Structure):
    _fields_ = [("temperature", c_float), ("beta", c_float), ("top_k", c_int32), ("top_p", c_float),
                ("use_final_max_rates", c_int32), ("method", c_int32), ("seed", c_uint64), ("first_global_jet", c_uint64)]

    def __init__(self, temperature=1.0, beta=0.075, top_k=0, top_p=0.0, use_final_max_rates=0, seed=0, first_global_jet=0,
                 method=0):
        # (`method` sits in what used to be padding of the C struct; positional construction keeps its historical order)
        super().__init__(temperature=temperature, beta=beta, top_k=top_k, top_p=top_p, use_final_max_rates=use_final_max_rates,
                         method=method, seed=seed, first_global_jet=first_global_jet)


_lib: Optional[ctypes.CDLL] = None

EXPORTS = [
    "mmf_abi_version", "mmf_last_error", "mmf_model_create", "mmf_model_destroy", "mmf_encoder_forward",
    "mmf_hybrid_step", "mmf_hybrid_step_status", "mmf_euler_step", "mmf_generate", "mmf_generate_n", "mmf_model_status",
    "mmf_generate_host", "mmf_launch_count", "mmf_jet_observables", "mmf_make_source",
    "mmf_sample_record_bytes", "mmf_pack_sample", "mmf_unpack_sample", "mmf_bridge_sample", "mmf_multitask_loss", "mmf_ema_update",
    "mmf_dbg_gemm", "mmf_dbg_gemm_resln", "mmf_dbg_gemm_qkv", "mmf_dbg_attention", "mmf_dbg_ring_plan",
    "mmf_profile_enable", "mmf_profile_num_classes", "mmf_profile_class_name", "mmf_profile_read",
]


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python multimodal-flows_b200/build.py` "
            "(there is no CPU or PyTorch fallback for the accelerated path)")
    L = ctypes.CDLL(LIB_PATH)
    L.mmf_abi_version.restype = c_int32
    L.mmf_last_error.restype = c_char_p
    L.mmf_model_create.argtypes = [POINTER(MmfModelDesc), POINTER(MmfWeightRef), c_int32, c_int32, POINTER(c_void_p)]
    L.mmf_model_destroy.argtypes = [c_void_p]
    L.mmf_model_destroy.restype = None
    L.mmf_launch_count.argtypes = [c_void_p]
    L.mmf_launch_count.restype = c_int64
    L.mmf_encoder_forward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                      c_void_p, c_void_p]
    L.mmf_hybrid_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, POINTER(MmfStepOptions),
                                  c_void_p, c_uint32, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]
    L.mmf_euler_step.argtypes = [c_void_p, c_void_p, c_float, c_int64, c_int32, c_void_p]
    L.mmf_jet_observables.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                      c_void_p, c_int32, c_void_p]
    L.mmf_make_source.argtypes = [c_void_p, c_int32, c_int32, c_int32, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int32, c_void_p]
    L.mmf_generate.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_float,
                               POINTER(MmfStepOptions), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.mmf_sample_record_bytes.argtypes = [c_int32]
    L.mmf_sample_record_bytes.restype = c_int64
    L.mmf_pack_sample.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p]
    L.mmf_unpack_sample.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    L.mmf_bridge_sample.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int32, c_void_p, c_void_p,
                                    c_uint64, c_uint64, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_void_p]
    L.mmf_multitask_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    L.mmf_ema_update.argtypes = [c_void_p, c_void_p, ctypes.c_double, c_int64, c_int32, c_void_p]
    L.mmf_generate_n.argtypes = L.mmf_generate.argtypes
    L.mmf_model_status.argtypes = [c_void_p, c_void_p]
    L.mmf_hybrid_step_status.argtypes = [c_int32, c_void_p]
    L.mmf_generate_host.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int32,
                                    c_float, POINTER(MmfStepOptions), c_void_p, c_void_p, c_void_p]
    L.mmf_dbg_gemm.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                               c_int32, c_void_p]
    L.mmf_dbg_gemm_resln.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                                     c_int32, c_void_p, c_void_p, c_int32, c_void_p]
    L.mmf_dbg_gemm_qkv.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                   c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]
    L.mmf_dbg_attention.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32,
                                    c_void_p, c_int32, c_void_p]
    L.mmf_profile_enable.argtypes = [c_void_p, c_int32]
    L.mmf_profile_num_classes.restype = c_int32
    L.mmf_profile_class_name.argtypes = [c_int32]
    L.mmf_profile_class_name.restype = c_char_p
    L.mmf_profile_read.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32]
    for name in EXPORTS:
        if name not in ("mmf_last_error", "mmf_model_destroy", "mmf_launch_count", "mmf_abi_version", "mmf_sample_record_bytes",
                        "mmf_profile_num_classes", "mmf_profile_class_name"):
            getattr(L, name).restype = c_int32
    if L.mmf_abi_version() != 2:
        raise RuntimeError("libmmf_b200.so ABI version mismatch")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().mmf_last_error()
        raise RuntimeError(f"libmmf_b200: {msg.decode() if msg else 'unknown error'} (status {rc})")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_handle(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def step_options(cfg, seed: int = 0, first_global_jet: int = 0, method: int = 0) -> MmfStepOptions:
    return MmfStepOptions(
        temperature=float(cfg.temperature if cfg.temperature is not None else 1.0), beta=float(cfg.beta),
        top_k=int(cfg.top_k or 0), top_p=float(cfg.top_p or 0.0),
        use_final_max_rates=int(bool(getattr(cfg, "use_final_max_rates", False))),
        method=int(method), seed=int(seed), first_global_jet=int(first_global_jet))


class NativeModel:
    """Owns one ``MmfModel*`` built from a reference-layout fp32 state_dict."""

    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], device: torch.device):
        L = lib()
        if device.type != "cuda":
            raise RuntimeError("the accelerated path runs on a CUDA device only (no CPU fallback)")
        self.device = device
        self.index = device.index if device.index is not None else torch.cuda.current_device()
        desc = MmfModelDesc(
            arch=ARCH[cfg.model], vocab_size=cfg.vocab_size, dim_continuous=cfg.dim_continuous, n_embd=cfg.n_embd,
            n_inner=cfg.n_inner if cfg.n_inner is not None else 4 * cfg.n_embd, n_head=cfg.n_head,
            n_layer=cfg.n_layer, n_layer_fused=getattr(cfg, "n_layer_fused", 0) or 0,
            n_embd_glob=getattr(cfg, "n_embd_glob", 0) or 0, qk_layernorm=int(bool(cfg.qk_layernorm)),
            max_num_particles=cfg.max_num_particles)
        keep = []
        refs = (MmfWeightRef * len(state_dict))()
        for i, (name, t) in enumerate(state_dict.items()):
            t = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(t)
            shape = (c_int64 * 4)(*(list(t.shape) + [1] * (4 - t.dim())))
            bname = name.encode()
            keep.append(bname)
            refs[i] = MmfWeightRef(bname, t.data_ptr(), t.dim(), shape)
        handle = c_void_p()
        check(L.mmf_model_create(ctypes.byref(desc), refs, len(state_dict), self.index, ctypes.byref(handle)))
        self.handle = handle
        self.vocab_size = cfg.vocab_size
        self.is_epic = cfg.model == "EPiC"
        self._pinned = {}                                 # cached pinned result buffers of generate_host, by shape

    def close(self) -> None:
        if getattr(self, "handle", None):
            lib().mmf_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(lib().mmf_launch_count(self.handle))

    def profile(self, on: bool) -> None:
        check(lib().mmf_profile_enable(self.handle, int(on)))

    def profile_read(self, reset: bool = True) -> Dict[str, Dict[str, float]]:
        """Per kernel class: accumulated ms, launches and algorithmic FLOPs since the last reset."""
        L = lib()
        n = L.mmf_profile_num_classes()
        ms = (ctypes.c_double * n)()
        cnt = (ctypes.c_int64 * n)()
        fl = (ctypes.c_double * n)()
        check(L.mmf_profile_read(self.handle, ms, cnt, fl, int(reset)))
        return {L.mmf_profile_class_name(i).decode(): {"ms": ms[i], "launches": int(cnt[i]), "flops": fl[i]}
                for i in range(n)}

    def forward(self, x, k, mask, t):
        """(B,D,3) f32, (B,D[,1]) i64, (B,D[,1]) i64, (B,) f32 -> vt (B,D,3), logits (B,D,V) (None for EPiC)."""
        B, D = x.shape[:2]
        x = x.contiguous().float()
        mask = mask.reshape(B, D).contiguous().long()
        t = t.contiguous().float()
        vt = torch.empty(B, D, 3, device=self.device, dtype=torch.float32)
        logits = None
        if not self.is_epic:
            k = k.reshape(B, D).contiguous().long()
            logits = torch.empty(B, D, self.vocab_size, device=self.device, dtype=torch.float32)
        check(lib().mmf_encoder_forward(self.handle, ptr(x), ptr(k) if not self.is_epic else None, ptr(mask), ptr(t),
                                        B, D, ptr(vt), ptr(logits), stream_handle(self.device)))
        return vt, logits

    def status(self) -> None:
        """Waits for the current stream and raises if a kernel of this handle met a token outside [0, V)
        (the reference's assert, model/MJB.py:177-182)."""
        check(lib().mmf_model_status(self.handle, stream_handle(self.device)))

    def generate(self, x0, k0, mask, t_grid, dt, opts: Optional[MmfStepOptions], u=None, forced_k=None,
                 want_rates=False, n_per_jet=None):
        """The N-step sampler on device tensors.  With ``n_per_jet`` (HOST int32 multiplicities of prefix masks,
        reference utils/aoj.py:882-883) the call is fully asynchronous (``mmf_generate_n``; errors via ``status()``);
        otherwise the device mask is read back to plan the tiles (``mmf_generate``, two host synchronisations)."""
        B, D = x0.shape[:2]
        N = int(t_grid.numel())
        x0 = x0.contiguous().float()
        if n_per_jet is not None:
            n_per_jet = torch.as_tensor(n_per_jet).to("cpu", torch.int32).contiguous()
            assert n_per_jet.numel() == B
        else:
            mask = mask.reshape(B, D).contiguous().long()
        tg = t_grid.detach().to("cpu", torch.float32).contiguous()
        x_out = torch.empty_like(x0)
        k_out = rates = None
        if not self.is_epic:
            k0 = k0.reshape(B, D).contiguous().long()
            k_out = torch.empty_like(k0)
            if want_rates:
                rates = torch.empty(B, D, self.vocab_size, device=self.device, dtype=torch.float32)
        if u is not None:
            u = u.contiguous().float()
            assert u.shape == (N, B, D, self.vocab_size)
        if forced_k is not None:
            forced_k = forced_k.reshape(N, B, D).contiguous().to(torch.uint8)
        fn = lib().mmf_generate if n_per_jet is None else lib().mmf_generate_n
        check(fn(self.handle, ptr(x0), ptr(k0) if not self.is_epic else None,
                 ptr(mask) if n_per_jet is None else n_per_jet.data_ptr(), B, D,
                 tg.data_ptr(), N, float(dt), ctypes.byref(opts) if opts is not None else None,
                 ptr(u), ptr(forced_k), ptr(x_out), ptr(k_out), ptr(rates), stream_handle(self.device)))
        return x_out, k_out, rates

    def generate_host(self, x0, k0, mask, t_grid, dt, opts: Optional[MmfStepOptions]):
        """Host tensors in, host tensors out (pinned or pageable); copies are inside the call."""
        B, D = x0.shape[:2]
        N = int(t_grid.numel())
        assert x0.device.type == "cpu" and mask.device.type == "cpu"
        x0 = x0.contiguous().float()
        mask = mask.reshape(B, D).contiguous().long()
        tg = t_grid.detach().to("cpu", torch.float32).contiguous()
        # pinned result buffers are cached per batch shape (pinning costs more than the copy); the caller gets clones
        key = (B, D)
        if key not in self._pinned:
            self._pinned[key] = (torch.empty(B, D, 3, dtype=torch.float32).pin_memory(),
                                 None if self.is_epic else torch.empty(B, D, dtype=torch.int64).pin_memory())
        px, pk = self._pinned[key]
        if not self.is_epic:
            k0 = k0.reshape(B, D).contiguous().long()
        check(lib().mmf_generate_host(self.handle, ptr(x0), ptr(k0) if not self.is_epic else None, ptr(mask), B, D,
                                      tg.data_ptr(), N, float(dt), ctypes.byref(opts) if opts is not None else None,
                                      ptr(px), ptr(pk), stream_handle(self.device)))
        return px.clone(), None if pk is None else pk.clone()


def hybrid_step(vt, logits, x, k, t, dt, opts: MmfStepOptions, u=None, step_index=0, want_rates=True):
    """In-place fused step on device tensors: x (B,D,3) f32, k (B,D) i64 (both contiguous)."""
    B, D = x.shape[:2]
    V = logits.shape[-1]
    assert x.is_cuda and x.is_contiguous() and k.is_contiguous() and k.dtype == torch.int64
    rates = torch.empty(B, D, V, device=x.device, dtype=torch.float32) if want_rates else None
    vt = vt.contiguous().float()
    logits = logits.contiguous().float()
    t = t.contiguous().float()
    if u is not None:
        u = u.contiguous().float()
    idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_hybrid_step(ptr(vt), ptr(logits), ptr(x), ptr(k), ptr(t), float(dt), ctypes.byref(opts), ptr(u),
                                int(step_index), B, D, V, ptr(rates), idx, stream_handle(x.device)))
    return rates


def hybrid_step_status(device) -> None:
    """Raises when ``hybrid_step`` met a token outside [0, V) on ``device`` since the last query (model/MJB.py:177-182)."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    check(lib().mmf_hybrid_step_status(idx, stream_handle(device)))


def make_source(mult_probs, num_jets, max_num_particles, vocab_size, seed, first_global_jet, device, discrete=True):
    """Source state on the device: returns (x0 (B,D,3) f32, k0 (B,D) i64 or None, mask (B,D) i64, n (B,) i32).
    mult_probs: D+1 non-negative weights of multiplicity 0..D (host)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("the source is built on the GPU (no CPU fallback)")
    B, D = int(num_jets), int(max_num_particles)
    probs = [float(v) for v in mult_probs]
    assert len(probs) == D + 1, "mult_probs needs one weight per multiplicity 0..D"
    probs_c = (ctypes.c_float * (D + 1))(*probs)
    x0 = torch.empty(B, D, 3, device=device, dtype=torch.float32)
    k0 = torch.empty(B, D, device=device, dtype=torch.int64) if discrete else None
    mask = torch.empty(B, D, device=device, dtype=torch.int64)
    n = torch.empty(B, device=device, dtype=torch.int32)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    check(lib().mmf_make_source(probs_c, B, D, int(vocab_size), int(seed), int(first_global_jet), ptr(x0), ptr(k0), ptr(mask), ptr(n),
                                idx, stream_handle(device)))
    return x0, k0, mask, n


OBS_COLUMNS = ("px", "py", "pz", "E", "pt", "m", "eta", "phi", "charge", "jet_charge", "multiplicity", "m2")


def jet_observables(x, k, mask, mean=None, std=None, vocab_size=9):
    """Fused jet observables of a sample on the device: returns (kin (B,12) f32 in OBS_COLUMNS order, counts (B,V) i32 or None).
    x (B,D,3) f32 standardised, k (B,D[,1]) i64 or None, mask (B,D[,1]) i64; mean / std: 3 floats each (None = identity)."""
    assert x.is_cuda and x.dtype == torch.float32
    B, D = x.shape[:2]
    x = x.contiguous()
    mask = mask.reshape(B, D).to(torch.int64).contiguous()
    if k is not None:
        k = k.reshape(B, D).to(torch.int64).contiguous()
    kin = torch.empty(B, len(OBS_COLUMNS), device=x.device, dtype=torch.float32)
    counts = torch.empty(B, vocab_size, device=x.device, dtype=torch.int32) if k is not None else None
    mean_c = (ctypes.c_float * 3)(*[float(v) for v in mean]) if mean is not None else None
    std_c = (ctypes.c_float * 3)(*[float(v) for v in std]) if std is not None else None
    idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_jet_observables(ptr(x), ptr(k), ptr(mask), mean_c, std_c, B, D, int(vocab_size), ptr(kin), ptr(counts), idx,
                                    stream_handle(x.device)))
    return kin, counts


def sample_record_bytes(D: int) -> int:
    return int(lib().mmf_sample_record_bytes(int(D)))


def pack_sample(x, k, mask, mean=None, std=None) -> torch.Tensor:
    """De-standardise + mask + narrow the generated sample into one record per jet: (B, R) uint8 on the device
    (reference utils/callbacks.py:52-57 fused with the int64 -> uint8 narrowing; layout in include/mmf_b200.h)."""
    assert x.is_cuda and x.dtype == torch.float32
    B, D = x.shape[:2]
    x = x.contiguous()
    mask = mask.reshape(B, D).to(torch.int64).contiguous()
    if k is not None:
        k = k.reshape(B, D).to(torch.int64).contiguous()
    rec = torch.empty(B, sample_record_bytes(D), device=x.device, dtype=torch.uint8)
    mean_c = (ctypes.c_float * 3)(*[float(v) for v in mean]) if mean is not None else None
    std_c = (ctypes.c_float * 3)(*[float(v) for v in std]) if std is not None else None
    idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_pack_sample(ptr(x), ptr(k), ptr(mask), mean_c, std_c, B, D, ptr(rec), idx, stream_handle(x.device)))
    return rec


def unpack_sample(rec, D: int, discrete: bool = True):
    """Inverse of ``pack_sample`` on the device: (x (B,D,3) f32, k (B,D) i64 or None, mask (B,D) i64)."""
    assert rec.is_cuda and rec.dtype == torch.uint8 and rec.is_contiguous() and rec.shape[1] == sample_record_bytes(D)
    B = rec.shape[0]
    x = torch.empty(B, D, 3, device=rec.device, dtype=torch.float32)
    k = torch.empty(B, D, device=rec.device, dtype=torch.int64) if discrete else None
    mask = torch.empty(B, D, device=rec.device, dtype=torch.int64)
    idx = rec.device.index if rec.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_unpack_sample(ptr(rec), B, D, ptr(x), ptr(k), ptr(mask), idx, stream_handle(rec.device)))
    return x, k, mask


def bridge_sample(x0, x1, k0, k1, t, sigma, beta, vocab_size, z=None, u=None, seed=0, first_global_jet=0):
    """xt (B,D,3), kt (B,D,1) of the two bridges at per-jet times t (reference model/CFM.py:171-184, model/MJB.py:197-257)."""
    assert x0.is_cuda
    B, D = x0.shape[:2]
    f = lambda a: a.contiguous().float()
    i = lambda a: a.reshape(B, D).contiguous().long()
    x0, x1, k0, k1, t = f(x0), f(x1), i(k0), i(k1), f(t)
    z = None if z is None else f(z)
    u = None if u is None else f(u)
    xt = torch.empty_like(x0)
    kt = torch.empty_like(k0)
    idx = x0.device.index if x0.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_bridge_sample(ptr(x0), ptr(x1), ptr(k0), ptr(k1), ptr(t), float(sigma), float(beta), int(vocab_size), ptr(z), ptr(u),
                                  int(seed), int(first_global_jet), B, D, ptr(xt), ptr(kt), idx, stream_handle(x0.device)))
    return xt, kt.unsqueeze(-1)


def multitask_loss(vt, logits, x0, x1, k1, mask, t, mode, n_embd=256, net=None):
    """(loss, loss_mse, loss_ce, w_mse, w_ce) as 0-dim tensors + per-jet (loss_mse, loss_ce) (reference model/MMF.py:152-168, 203-233).
    mode "sum" | "time-weighted"; net = (c_fc.weight, c_fc.bias, c_proj.weight, c_proj.bias) of loss_combine.uncertainty_net."""
    B, D = x0.shape[:2]
    V = logits.shape[-1]
    f = lambda a: a.contiguous().float()
    vt, logits, x0, x1, t = f(vt), f(logits), f(x0), f(x1), f(t)
    k1 = k1.reshape(B, D).contiguous().long()
    mask = mask.reshape(B, D).contiguous().long()
    per_jet = torch.empty(2, B, device=x0.device, dtype=torch.float32)
    out = torch.empty(5, device=x0.device, dtype=torch.float32)
    m = {"sum": 0, "time-weighted": 1}[mode]
    ws = [None] * 4 if net is None else [f(w) for w in net]
    idx = x0.device.index if x0.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_multitask_loss(ptr(vt), ptr(logits), ptr(x0), ptr(x1), ptr(k1), ptr(mask), ptr(t), B, D, V, m, int(n_embd),
                                   ptr(ws[0]), ptr(ws[1]), ptr(ws[2]), ptr(ws[3]), ptr(per_jet), ptr(out), idx, stream_handle(x0.device)))
    return out, per_jet


def ema_update(ema: torch.Tensor, p: torch.Tensor, decay: float) -> None:
    """ema <- decay * ema + (1 - decay) * p in place (fp32, device), timm ModelEmaV2.update for one tensor."""
    assert ema.is_cuda and ema.dtype == torch.float32 and ema.is_contiguous() and p.shape == ema.shape
    p = p.detach().to(ema.device, torch.float32).contiguous()
    idx = ema.device.index if ema.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_ema_update(ptr(ema), ptr(p), float(decay), ema.numel(), idx, stream_handle(ema.device)))


def euler_step(vt, x, dt):
    assert x.is_cuda and x.is_contiguous()
    vt = vt.contiguous().float()
    idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
    check(lib().mmf_euler_step(ptr(vt), ptr(x), float(dt), x.numel(), idx, stream_handle(x.device)))
