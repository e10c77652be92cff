"""ctypes binding of ``oracle/step_oracle.c`` (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmmf_step_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "step_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libmmf_step_oracle.so"])
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        f = ctypes.c_float
        p = ctypes.c_void_p
        _lib.mmf_det_expf.restype = f
        _lib.mmf_det_expf.argtypes = [f]
        _lib.mmf_oracle_thermostat.argtypes = [f, f, ctypes.c_int, p, p]
        _lib.mmf_oracle_hybrid_step.restype = ctypes.c_int
        _lib.mmf_oracle_hybrid_step.argtypes = [p, p, p, p, p, f, f, f, ctypes.c_int, ctypes.c_int, f, p,
                                                ctypes.c_int, ctypes.c_int, p, p, p]
        _lib.mmf_oracle_euler_step.restype = ctypes.c_int
        _lib.mmf_oracle_euler_step.argtypes = [p, p, p, p, p, f, f, ctypes.c_int, ctypes.c_int, f, p,
                                               ctypes.c_int, ctypes.c_int, p, p, p]
    return _lib


def det_expf(x: float) -> float:
    return float(lib().mmf_det_expf(float(x)))


def thermostat(t: float, beta: float, V: int):
    w = ctypes.c_float()
    c = ctypes.c_float()
    lib().mmf_oracle_thermostat(float(t), float(beta), int(V), ctypes.byref(w), ctypes.byref(c))
    return w.value, c.value


def hybrid_step(vt, logits, x, k, t, dt, u, *, temperature=1.0, beta=0.075, vocab_size=9,
                top_k=None, top_p=None, want_rates=True):
    """Same contract as ``mmf_oracle.hybrid_step`` (k is (B,D,1) int64) but with the
    deterministic arithmetic of step_oracle.c.  Returns (x', k', rates)."""
    B, D = x.shape[:2]
    V = vocab_size
    c = lambda a, dt_: np.ascontiguousarray(a.detach().cpu().numpy().astype(dt_))
    vt_, lg_, x_, u_, t_ = c(vt, np.float32), c(logits, np.float32), c(x, np.float32), c(u, np.float32), c(t, np.float32)
    k_ = c(k.reshape(B, D), np.int64)
    xo = np.empty_like(x_)
    ko = np.empty_like(k_)
    ro = np.empty((B, D, V), np.float32) if want_rates else None
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    bad = lib().mmf_oracle_hybrid_step(ptr(vt_), ptr(lg_), ptr(x_), ptr(k_), ptr(t_), float(dt), float(temperature),
                                       float(beta), V, int(top_k or 0), float(top_p or 0.0), ptr(u_), B, D,
                                       ptr(xo), ptr(ko), ptr(ro))
    if bad:
        raise AssertionError(f"{bad} tokens outside [0,{V})")
    return (torch.from_numpy(xo), torch.from_numpy(ko).unsqueeze(-1),
            torch.from_numpy(ro) if want_rates else None)


def euler_categorical_step(vt, logits, x, k, t, dt, u, *, beta=0.075, vocab_size=9, top_k=None, top_p=None,
                           want_rates=True):
    """``HybridSolver.euler_step`` (reference model/solvers.py:62-91, T = 1) with the deterministic arithmetic of
    step_oracle.c; ``u`` is (B,D) - one uniform per particle.  Returns (x', k' (B,D,1), rates)."""
    B, D = x.shape[:2]
    V = vocab_size
    c = lambda a, dt_: np.ascontiguousarray(a.detach().cpu().numpy().astype(dt_))
    vt_, lg_, x_, u_, t_ = c(vt, np.float32), c(logits, np.float32), c(x, np.float32), c(u, np.float32), c(t, np.float32)
    k_ = c(k.reshape(B, D), np.int64)
    xo = np.empty_like(x_)
    ko = np.empty_like(k_)
    ro = np.empty((B, D, V), np.float32) if want_rates else None
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    bad = lib().mmf_oracle_euler_step(ptr(vt_), ptr(lg_), ptr(x_), ptr(k_), ptr(t_), float(dt), float(beta), V,
                                      int(top_k or 0), float(top_p or 0.0), ptr(u_), B, D, ptr(xo), ptr(ko), ptr(ro))
    if bad:
        raise AssertionError(f"{bad} tokens outside [0,{V})")
    return (torch.from_numpy(xo), torch.from_numpy(ko).unsqueeze(-1), torch.from_numpy(ro) if want_rates else None)
