"""ctypes signatures of the training-step operators (``include/mmf_b200_train.h``) on top of ``_abi.lib()``.

Thin wrappers over torch tensors: a 2-D tensor (or a column slice of one) crosses the boundary as (data_ptr, row pitch).
"""
from __future__ import annotations

from ctypes import c_float, c_int32, c_int64, c_void_p
from typing import Optional

import torch

from . import _abi

P, I32, I64, F = c_void_p, c_int32, c_int64, c_float

SIGNATURES = {
    "mmf_tr_gemm": [P, I64, P, I64, P, I64, I32, I32, I32, P, I32, I32, P, I64, P, I64, P, I64, P, P],
    "mmf_tr_gemm_qkv": [P, I64, P, I64, P, P, I64, P, I64, I32, I32, I32, I32, P, P, P, P, P],
    "mmf_tr_gemm_tn": [P, I64, P, I64, P, I64, I32, I32, I32, I32, P],
    "mmf_tr_sgemm": [P, I64, I64, P, I64, I64, P, I64, I32, I32, I32, P, I32, P],
    "mmf_tr_cast_transpose": [P, I64, I32, I32, I32, P, I64, P, I64, P, P],
    "mmf_tr_weights_transpose": [P, P, P, I32, I32, P],
    "mmf_tr_pack": [P, P, P, P, P, P, I32, I32, P, P, P, P, P, P],
    "mmf_tr_time_embed": [P, P, I32, I32, I32, P, I64, P],
    "mmf_tr_embed_x_fwd": [P, I32, P, P, I32, P, I64, P],
    "mmf_tr_embed_x_bwd": [P, I64, P, I32, P, P, I32, P, P, P],
    "mmf_tr_embed_y_fwd": [P, I32, P, I32, I32, P, I64, P],
    "mmf_tr_embed_y_bwd": [P, I64, P, I32, P, I32, I32, P, P],
    "mmf_tr_ln_fwd": [P, I64, P, I64, P, P, P, I64, P, I32, I32, P, I64, P, I64, P, P, P],
    "mmf_tr_ln_bwd": [P, I64, P, I64, P, I64, P, P, P, I32, I32, P, I64, I32, P, P, P, I64, P, P],
    "mmf_tr_qkln_fwd": [P, I64, I32, I32, I32, P, P, P, P, P, P, I64, P],
    "mmf_tr_qkln_bwd": [P, I64, P, I64, I32, I32, I32, P, P, P, P, P, P, P],
    "mmf_tr_attn_fwd": [P, I64, P, I64, P, I64, P, P, I32, I32, I32, I32, I32, P, I64, P, P],
    "mmf_tr_attn_bwd": [P, I64, P, I64, P, P, I64, P, I64, P, I64, P, P, I32, I32, I32, I32, I32, P, I64, I32, P],
    "mmf_tr_attn_tc_fwd": [P, I64, P, I64, P, I64, I32, I32, I32, P, P, I32, P, P, P, P, I64, P],
    "mmf_tr_attn_tc_bwd": [P, I64, P, I64, P, I64, P, I64, I32, I32, I32, P, P, I32, P, P, P, P, I64, P],
    "mmf_tr_gelu_fwd": [P, P, I64, I32, P],
    "mmf_tr_gelu_bwd": [P, P, P, I64, I32, P],
    "mmf_tr_add": [P, I64, P, I64, P, I64, P, I64, P, I32, I32, P],
    "mmf_tr_jet_sum": [P, I64, P, I32, I32, P, I64, I32, P],
    "mmf_tr_head_fwd": [P, I64, I32, P, P, P, P, I32, I32, P, P, P],
    "mmf_tr_head_bwd": [P, P, P, P, I64, I32, P, P, I32, I32, P, P, P, P, P, P],
    "mmf_tr_loss_fwd": [P, P, P, P, P, I32, I32, P, P, P],
    "mmf_tr_loss_combine": [P, P, P, I32, P, P, P, P, P],
    "mmf_tr_loss_bwd": [P, P, P, P, P, P, P, P, I32, I32, I32, P, P, P],
    "mmf_tr_sumsq": [P, I64, P, P],
    "mmf_tr_adam": [P, P, P, P, I64, F, F, F, F, I32, P, F, F, P, P],
}
TRAIN_EXPORTS = sorted(SIGNATURES)

_bound = None


def lib():
    global _bound
    if _bound is None:
        L = _abi.lib()
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = c_int32
        _bound = L
    return _bound


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _ld(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else (t.stride(0) if t.dim() > 1 else t.shape[0])


class Ops:
    """The operators bound to one CUDA stream (torch's current stream of ``device`` at call time)."""

    def __init__(self, device: torch.device):
        self.L = lib()
        self.device = device
        self.launches = 0

    def _s(self):
        self.launches += 1
        return torch.cuda.current_stream(self.device).cuda_stream

    # C[M,N] (+)= A[M,K] B[N,K]^T (+ bias)
    def gemm(self, A, B, C, bias=None, mode=0, ksplit=1, aux=None, resid=None, tadd=None, row_jet=None):
        M, K = A.shape
        N = B.shape[0]
        assert B.shape[1] == K and tuple(C.shape) == (M, N) and A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
        assert (mode in (3, 4)) == (aux is not None) and (aux is None or (tuple(aux.shape) == (M, N) and aux.stride(1) == 1))
        assert (mode == 5) == (resid is not None) and (resid is None or tuple(resid.shape) == (M, N))
        _abi.check(self.L.mmf_tr_gemm(_p(A), A.stride(0), _p(B), B.stride(0), _p(C), C.stride(0), M, N, K, _p(bias), mode, ksplit, _p(aux),
                                      _ld(aux), _p(resid), _ld(resid), _p(tadd), _ld(tadd), _p(row_jet), self._s()))

    # qkv = A W^T + b [M,3C]; qkn = per-head LayerNorm of its q | k sections [M,2C]
    def gemm_qkv(self, A, W, bias, qkv, qkn, H, qg, qb, kg, kb):
        M, K = A.shape
        C = W.shape[0] // 3
        assert tuple(qkv.shape) == (M, 3 * C) and tuple(qkn.shape) == (M, 2 * C) and W.shape[1] == K
        _abi.check(self.L.mmf_tr_gemm_qkv(_p(A), A.stride(0), _p(W), W.stride(0), _p(bias), _p(qkv), qkv.stride(0), _p(qkn), qkn.stride(0),
                                          M, C, K, H, _p(qg), _p(qb), _p(kg), _p(kb), self._s()))

    # C[M,N] += A^T B, A [K,M], B [K,N] row-major bf16
    def gemm_tn(self, A, B, C, ksplit=1):
        K, M = A.shape
        N = B.shape[1]
        assert B.shape[0] == K and tuple(C.shape) == (M, N) and A.stride(1) == 1 and B.stride(1) == 1 and C.stride(1) == 1
        _abi.check(self.L.mmf_tr_gemm_tn(_p(A), A.stride(0), _p(B), B.stride(0), _p(C), C.stride(0), M, N, K, ksplit, self._s()))

    def sgemm(self, A, sam, sak, B, sbk, sbn, C, M, N, K, bias=None, accumulate=False):
        _abi.check(self.L.mmf_tr_sgemm(_p(A), sam, sak, _p(B), sbk, sbn, _p(C), C.stride(0) if C.dim() > 1 else N, M, N, K, _p(bias),
                                       int(accumulate), self._s()))

    def cast_transpose(self, x, out=None, outT=None, colsum=None):
        rows, cols = x.shape
        _abi.check(self.L.mmf_tr_cast_transpose(_p(x), x.stride(0), int(x.dtype == torch.float32), rows, cols, _p(out), _ld(out), _p(outT),
                                                _ld(outT), _p(colsum), self._s()))

    def weights_transpose(self, params, paramsT, jobs, n_jobs, n_tiles):
        _abi.check(self.L.mmf_tr_weights_transpose(_p(params), _p(paramsT), _p(jobs), n_jobs, n_tiles, self._s()))

    def pack(self, xt, kt, x0, x1, k1, row_slot, V, xs, ks, tgt, k1p, err):
        _abi.check(self.L.mmf_tr_pack(_p(xt), _p(kt), _p(x0), _p(x1), _p(k1), _p(row_slot), row_slot.shape[0], V, _p(xs), _p(ks), _p(tgt),
                                      _p(k1p), _p(err), self._s()))

    def time_embed(self, t, dim, dup, out, perm=None):
        _abi.check(self.L.mmf_tr_time_embed(_p(t), _p(perm), t.shape[0], dim, int(dup), _p(out), out.stride(0), self._s()))

    def embed_x_fwd(self, xs, w0, b0, h):
        _abi.check(self.L.mmf_tr_embed_x_fwd(_p(xs), xs.shape[0], _p(w0), _p(b0), w0.shape[0], _p(h), h.stride(0), self._s()))

    def embed_x_bwd(self, dh, xs, w0, b0, dw0, db0):
        _abi.check(self.L.mmf_tr_embed_x_bwd(_p(dh), dh.stride(0), _p(xs), xs.shape[0], _p(w0), _p(b0), w0.shape[0], _p(dw0), _p(db0), self._s()))

    def embed_y_fwd(self, ks, emb, g):
        _abi.check(self.L.mmf_tr_embed_y_fwd(_p(ks), ks.shape[0], _p(emb), emb.shape[1], emb.shape[0], _p(g), g.stride(0), self._s()))

    def embed_y_bwd(self, dg, ks, emb, demb):
        _abi.check(self.L.mmf_tr_embed_y_bwd(_p(dg), dg.stride(0), _p(ks), ks.shape[0], _p(emb), emb.shape[1], emb.shape[0], _p(demb), self._s()))

    def ln_fwd(self, x, g, b, mean, rstd, add=None, tadd=None, row_jet=None, out16=None, out32=None):
        M, C = x.shape
        _abi.check(self.L.mmf_tr_ln_fwd(_p(x), x.stride(0), _p(add), _ld(add), _p(g), _p(b), _p(tadd), _ld(tadd), _p(row_jet), M, C,
                                        _p(out16), _ld(out16), _p(out32), _ld(out32), _p(mean), _p(rstd), self._s()))

    def ln_bwd(self, dy, x, mean, rstd, g, dx, dg, db, add=None, accumulate=False, dx16=None, dxsum=None):
        M, C = x.shape
        _abi.check(self.L.mmf_tr_ln_bwd(_p(dy), dy.stride(0), _p(x), x.stride(0), _p(add), _ld(add), _p(mean), _p(rstd), _p(g), M, C,
                                        _p(dx), dx.stride(0), int(accumulate), _p(dg), _p(db), _p(dx16), _ld(dx16), _p(dxsum), self._s()))

    def qkln_fwd(self, qkv, C, H, qg, qb, kg, kb, qn, kn):
        _abi.check(self.L.mmf_tr_qkln_fwd(_p(qkv), qkv.stride(0), qkv.shape[0], C, H, _p(qg), _p(qb), _p(kg), _p(kb), _p(qn), _p(kn),
                                          qn.stride(0), self._s()))

    def qkln_bwd(self, dqkv, qkv, C, H, qg, kg, dqg, dqb, dkg, dkb):
        _abi.check(self.L.mmf_tr_qkln_bwd(_p(dqkv), dqkv.stride(0), _p(qkv), qkv.stride(0), qkv.shape[0], C, H, _p(qg), _p(kg), _p(dqg),
                                          _p(dqb), _p(dkg), _p(dkb), self._s()))

    def attn_fwd(self, qn, kn, v, jet_off, p_off, B, H, hs, nmax, o, P_, min_n=0):
        _abi.check(self.L.mmf_tr_attn_fwd(_p(qn), qn.stride(0), _p(kn), kn.stride(0), _p(v), v.stride(0), _p(jet_off), _p(p_off), B, H, hs,
                                          nmax, min_n, _p(o), o.stride(0), _p(P_), self._s()))

    def attn_bwd(self, dO, o, P_, qn, kn, v, jet_off, p_off, B, H, hs, nmax, dqkv, C, min_n=0):
        _abi.check(self.L.mmf_tr_attn_bwd(_p(dO), dO.stride(0), _p(o), o.stride(0), _p(P_), _p(qn), qn.stride(0), _p(kn), kn.stride(0),
                                          _p(v), v.stride(0), _p(jet_off), _p(p_off), B, H, hs, nmax, min_n, _p(dqkv), dqkv.stride(0), C,
                                          self._s()))

    # tensor-core attention over items of whole jets (<= 128 rows each)
    def attn_tc_fwd(self, qn, kn, v, hs, items, n_items, grid_items, row_jet, jet_off, stats, o):
        M, C = qn.shape
        _abi.check(self.L.mmf_tr_attn_tc_fwd(_p(qn), qn.stride(0), _p(kn), kn.stride(0), _p(v), v.stride(0), M, C, hs, _p(items), _p(n_items),
                                             grid_items, _p(row_jet), _p(jet_off), _p(stats), _p(o), o.stride(0), self._s()))

    def attn_tc_bwd(self, dO, qn, kn, v, hs, items, n_items, grid_items, row_jet, jet_off, stats, dqkv):
        M, C = qn.shape
        _abi.check(self.L.mmf_tr_attn_tc_bwd(_p(dO), dO.stride(0), _p(qn), qn.stride(0), _p(kn), kn.stride(0), _p(v), v.stride(0), M, C, hs,
                                             _p(items), _p(n_items), grid_items, _p(row_jet), _p(jet_off), _p(stats), _p(dqkv),
                                             dqkv.stride(0), self._s()))

    def gelu_fwd(self, z, h):
        assert z.is_contiguous() and h.is_contiguous()
        _abi.check(self.L.mmf_tr_gelu_fwd(_p(z), _p(h), z.numel(), int(z.dtype == torch.float32), self._s()))

    def gelu_bwd(self, dh, z, dz):
        assert z.is_contiguous() and dh.is_contiguous() and dz.is_contiguous()
        _abi.check(self.L.mmf_tr_gelu_bwd(_p(dh), _p(z), _p(dz), z.numel(), int(z.dtype == torch.float32), self._s()))

    def add(self, out, a, y=None, tadd=None, row_jet=None):
        M, C = a.shape
        _abi.check(self.L.mmf_tr_add(_p(out), out.stride(0), _p(a), a.stride(0), _p(y), _ld(y), _p(tadd), _ld(tadd), _p(row_jet), M, C, self._s()))

    def jet_sum(self, g, jet_off, B, out, accumulate=False):
        _abi.check(self.L.mmf_tr_jet_sum(_p(g), g.stride(0), _p(jet_off), B, g.shape[1], _p(out), out.stride(0), int(accumulate), self._s()))

    def head_fwd(self, h, I, wx, bx, wy, by, vt, logits):
        _abi.check(self.L.mmf_tr_head_fwd(_p(h), h.stride(0), I, _p(wx), _p(bx), _p(wy), _p(by), wy.shape[0], h.shape[0], _p(vt), _p(logits), self._s()))

    def head_bwd(self, dvt, dlog, h, z, I, wx, wy, dz, dwx, dbx, dwy, dby):
        _abi.check(self.L.mmf_tr_head_bwd(_p(dvt), _p(dlog), _p(h), _p(z), h.stride(0), I, _p(wx), _p(wy), wy.shape[0], h.shape[0], _p(dz),
                                          _p(dwx), _p(dbx), _p(dwy), _p(dby), self._s()))

    def loss_fwd(self, vt, logits, tgt, k1, jet_off, B, V, loss_mse, loss_ce):
        _abi.check(self.L.mmf_tr_loss_fwd(_p(vt), _p(logits), _p(tgt), _p(k1), _p(jet_off), B, V, _p(loss_mse), _p(loss_ce), self._s()))

    def loss_combine(self, loss_mse, loss_ce, u, out5, gl1, gl2, du):
        _abi.check(self.L.mmf_tr_loss_combine(_p(loss_mse), _p(loss_ce), _p(u), loss_mse.shape[0], _p(out5), _p(gl1), _p(gl2), _p(du), self._s()))

    def loss_bwd(self, vt, logits, tgt, k1, row_jet, jet_off, gl1, gl2, V, dvt, dlog):
        _abi.check(self.L.mmf_tr_loss_bwd(_p(vt), _p(logits), _p(tgt), _p(k1), _p(row_jet), _p(jet_off), _p(gl1), _p(gl2), vt.shape[0],
                                          gl1.shape[0], V,
                                          _p(dvt), _p(dlog), self._s()))

    def sumsq(self, g, out):
        _abi.check(self.L.mmf_tr_sumsq(_p(g), g.numel(), _p(out), self._s()))

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, step, sumsq=None, max_norm=0.0, grad_scale=1.0, p16=None):
        _abi.check(self.L.mmf_tr_adam(_p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, step, _p(sumsq), max_norm, grad_scale,
                                      _p(p16), self._s()))
