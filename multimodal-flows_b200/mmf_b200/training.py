"""The MMF training step on the device (BASELINE config #5): bf16 tensor-core forward and backward of
``MultiModalFlowBridge.loss`` (reference model/MMF.py:138-170) for ParticleFormer / FusedParticleFormer, Adam
(model/MMF.py:77-78), Lightning's ``gradient_clip_val=1.0`` (scripts/train_mmf.py:166) and the DDP gradient average
(``strategy='ddp'``, :163) as ONE all-reduce of the flat gradient buffer.

The arithmetic lives in ``libmmf_b200.so`` (``include/mmf_b200_train.h``); this file is the host-side sequencing the
reference leaves to torch autograd: which operator runs on which buffer, in which order.  There is no torch autograd,
no torch math on the path and no CPU fallback.

Parameters stay ``nn.Parameter`` objects with the reference's ``state_dict`` keys; ``TrainEngine`` re-homes their storage in
one flat fp32 buffer (``.data`` becomes a view), with flat buffers for the gradients (``.grad`` views), Adam moments and the
bf16 (and transposed bf16) operand copies the GEMMs read.

Layout: packed rows (one row per real particle, jets contiguous), see ``mmf_b200_train.h``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import _abi
from ._train_abi import Ops
from .tensorclass import DataCoupling, TensorMultiModal

_ALIGN = 64                                      # elements: every parameter starts on a 256-byte boundary (TMA needs 16)


class _Plan:
    """Row maps of one batch, built on the host from the mask (any mask works; the reference's are prefix masks).

    The packed order of the jets is the planner's choice (nothing downstream depends on it: the step returns batch means and
    parameter gradients): jets of at most 128 particles are packed best-fit-decreasing into tiles of at most 128 rows - the work
    items of the tensor-core attention kernels - and laid out tile after tile; jets of 129 ... 150 particles follow (CUDA-core
    attention kernels, their probabilities are kept: p_off counts those jets only), empty jets last.  perm[p] = index in the batch
    of packed jet p (the per-jet time is read through it)."""
    padded = False

    def __init__(self, mask: torch.Tensor, device: torch.device, upload: bool = True):
        import bisect
        m = mask.detach().reshape(mask.shape[0], -1).to("cpu").numpy() != 0
        self.B, self.D = m.shape
        n = m.sum(1).astype(np.int64)
        self.n = n
        self.M = int(n.sum())
        self.nmax = int(n.max()) if self.B else 0
        self.Mp = (self.M + 63) // 64 * 64
        bins, free = [], []                                   # free: sorted (remaining rows, bin)
        for b in sorted((int(b) for b in np.flatnonzero((n > 0) & (n <= 128))), key=lambda b: -int(n[b])):
            nb = int(n[b])
            i = bisect.bisect_left(free, (nb, -1))
            if i < len(free):
                rem, bi = free.pop(i)
                bins[bi].append(b)
            else:
                rem, bi = 128, len(bins)
                bins.append([b])
            bisect.insort(free, (rem - nb, bi))
        order = [b for jets in bins for b in jets] + [int(b) for b in np.flatnonzero(n > 128)] + [int(b) for b in np.flatnonzero(n == 0)]
        self.h_perm = np.asarray(order, np.int32)
        npk = n[self.h_perm] if self.B else n
        self.n_packed = npk
        self.h_jet_off = np.zeros(self.B + 1, np.int32)
        np.cumsum(npk, out=self.h_jet_off[1:])
        big = npk > 128
        self.has_big = bool(big.any())
        self.h_p_off = np.zeros(self.B + 1, np.int64)
        np.cumsum(np.where(big, npk * npk, 0), out=self.h_p_off[1:])
        self.sum_n2 = int(self.h_p_off[-1])
        items, row = [], 0
        for jets in bins:
            rows = int(sum(int(n[b]) for b in jets))
            items.append((row, rows))
            row += rows
        self.h_items = np.asarray(items, np.int32).reshape(-1, 2)
        self.grid_items = len(items)
        pm = m[self.h_perm] if self.B else m                  # masks in packed jet order
        pj, pd = np.nonzero(pm)
        self.h_row_slot = (self.h_perm[pj].astype(np.int64) * self.D + pd).astype(np.int32)
        self.h_row_jet = pj.astype(np.int32)
        up = lambda a: torch.from_numpy(a).to(device, non_blocking=True)
        self.row_slot = up(self.h_row_slot)
        if upload:
            self.jet_off, self.p_off, self.row_jet, self.perm = up(self.h_jet_off), up(self.h_p_off), up(self.h_row_jet), up(self.h_perm)
            self.items = up(self.h_items if len(items) else np.zeros((1, 2), np.int32))
            self.n_items = up(np.asarray([len(items)], np.int32))


class _GraphSlot:
    """Static buffers of one captured forward + backward program: a batch of B jets padded to `rows` packed rows.  Rows at and
    beyond the last jet belong to no jet: attention never touches them, the loss gives them zero gradient, so they add
    nothing to any parameter gradient - the same CUDA graph serves every batch of B jets with at most `rows` particles."""
    padded = True

    def __init__(self, B: int, D: int, rows: int, has_big: bool, device: torch.device):
        # has_big: the graph holds the CUDA-core attention launches for jets of more than 128 particles (and their probability
        # buffers, sized for the worst case); batches without such jets replay a graph without them
        self.B, self.D, self.M, self.Mp, self.nmax, self.sum_n2 = B, D, rows, rows, D, (B * D * D if has_big else 0)
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device=device, dtype=dt)
        self.jet_off, self.p_off, self.row_jet = z(B + 1, dt=torch.int32), z(B + 1, dt=torch.int64), z(rows, dt=torch.int32)
        self.xs, self.tg, self.ks, self.k1p, self.t = z(rows, 3), z(rows, 3), z(rows, dt=torch.int32), z(rows, dt=torch.int32), z(B)
        self.items, self.n_items, self.grid_items, self.has_big = z(B, 2, dt=torch.int32), z(1, dt=torch.int32), B, bool(has_big)
        self.perm = z(B, dt=torch.int32)
        self.pin = [torch.zeros(B + 1, dtype=torch.int32).pin_memory(), torch.zeros(B + 1, dtype=torch.int64).pin_memory(),
                    torch.zeros(rows, dtype=torch.int32).pin_memory(), torch.zeros(B, 2, dtype=torch.int32).pin_memory(),
                    torch.zeros(1, dtype=torch.int32).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()]
        self.graph, self.out5, self.busy = None, None, None

    def load(self, plan: _Plan):
        if self.busy is not None:
            self.busy.synchronize()                            # the previous upload still reads the pinned staging
        self.pin[0].copy_(torch.from_numpy(plan.h_jet_off))
        self.pin[1].copy_(torch.from_numpy(plan.h_p_off))
        self.pin[2][: plan.M].copy_(torch.from_numpy(plan.h_row_jet))
        self.pin[3][: plan.grid_items].copy_(torch.from_numpy(plan.h_items))
        self.pin[4][0] = plan.grid_items
        self.pin[5].copy_(torch.from_numpy(plan.h_perm))
        self.perm.copy_(self.pin[5], non_blocking=True)
        self.items.copy_(self.pin[3], non_blocking=True)
        self.n_items.copy_(self.pin[4], non_blocking=True)
        self.jet_off.copy_(self.pin[0], non_blocking=True)
        self.p_off.copy_(self.pin[1], non_blocking=True)
        self.row_jet[: plan.M].copy_(self.pin[2][: plan.M], non_blocking=True)
        self.busy = torch.cuda.Event()
        self.busy.record()


class TrainEngine:
    """fwd + bwd + optimiser for one ``MultiModalFlowBridge`` (its encoder and its ``MultiTaskLoss``)."""

    def __init__(self, module, lr: Optional[float] = None, betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 1.0,
                 use_graphs: bool = False, _ops=None):
        """``_ops`` is a test seam (tests/mock_train_ops.py checks the sequencing below on the CPU); the product path has no
        default for it other than the library."""
        cfg = module.config
        if cfg.model not in ("ParticleFormer", "FusedParticleFormer"):
            raise NotImplementedError("the training step covers ParticleFormer and FusedParticleFormer")
        if cfg.multitask_loss not in ("sum", "weighted", "time-weighted"):
            raise ValueError(f"unknown multitask_loss '{cfg.multitask_loss}'")
        if getattr(cfg, "dropout", 0.0):
            raise NotImplementedError("dropout > 0 is not supported (the reference trains with dropout 0.0, scripts/train_mmf.py:55)")
        dev = next(module.parameters()).device
        if _ops is None and dev.type != "cuda":
            raise RuntimeError("the training step runs through libmmf_b200.so on a CUDA device (there is no CPU fallback)")
        self.module, self.cfg, self.device = module, cfg, dev
        self.model_ref = module.model                        # the encoder whose parameters live in the flat buffers
        self.ops = _ops if _ops is not None else Ops(dev)
        self.lr = float(lr if lr is not None else getattr(cfg, "lr", 1e-3))
        self.betas, self.eps, self.max_norm = betas, eps, float(max_norm)
        self.step_count = 0
        self.pf = cfg.model == "ParticleFormer"
        self.E, self.h, self.I, self.H, self.V = cfg.n_embd, cfg.n_embd // 2, cfg.n_inner or 4 * cfg.n_embd, cfg.n_head, cfg.vocab_size
        if self.E != 256:
            raise NotImplementedError("the training kernels are instantiated for n_embd = 256")

        named = [("model." + n, p) for n, p in module.model.named_parameters()]
        named += [("loss_combine." + n, p) for n, p in module.loss_combine.named_parameters()]
        self.names = [n for n, _ in named]
        self.off: Dict[str, int] = {}
        self.shape: Dict[str, tuple] = {}
        total = 0
        for n, p in named:
            self.off[n], self.shape[n] = total, tuple(p.shape)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = total
        self.P = torch.zeros(total, device=dev)
        self.G = torch.zeros(total, device=dev)
        self.m = torch.zeros(total, device=dev)
        self.v = torch.zeros(total, device=dev)
        self.P16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        self.PT16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        with torch.no_grad():
            for n, p in named:
                view = self.P[self.off[n]: self.off[n] + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view                                             # the module now lives in the flat buffer
                p.grad = self.G[self.off[n]: self.off[n] + p.numel()].view(p.shape)
        # two-dimensional weights that feed the tensor-core GEMMs: transposed bf16 copies for the data-gradient products
        jobs, tile0 = [], 0
        for n in self.names:
            s = self.shape[n]
            if n.startswith("model.") and len(s) == 2 and s[0] % 8 == 0 and s[1] % 64 == 0 and "time_expand" not in n:
                jobs.append((self.off[n], self.off[n], s[0], s[1], tile0, 0))
                tile0 += ((s[0] + 31) // 32) * ((s[1] + 31) // 32)
        rec = np.zeros(len(jobs), dtype=[("src", "<i8"), ("dst", "<i8"), ("rows", "<i4"), ("cols", "<i4"), ("tile0", "<i4"), ("pad", "<i4")])
        for i, j in enumerate(jobs):
            rec[i] = j
        self._jobs = torch.from_numpy(rec.view(np.uint8).copy()).to(dev)
        self._n_jobs, self._n_tiles = len(jobs), tile0
        self._sumsq, self._ones = torch.zeros(2048, device=dev), None       # [0] the squared gradient norm, the rest scratch
        self._err = torch.zeros(1, device=dev, dtype=torch.int32)
        if _ops is None and torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            # what DistributedDataParallel does at construction: every replica starts from rank 0's parameters (the randomly
            # initialised uncertainty net of MultiTaskLoss differs from process to process otherwise)
            torch.distributed.broadcast(self.P, src=0)
        self.refresh_operands()
        self.use_graphs, self.graph_rows, self._slots, self.max_graphs = bool(use_graphs) and _ops is None, 1024, {}, 6
        self.concurrent, self._streams, self._wstreams, self._wused, self._keep = _ops is None, [], {}, set(), []
        self.last_plan = None

    # ---- parameter views -------------------------------------------------------------------------------------------------
    def _v(self, buf, name):
        s = self.shape[name]
        n = int(np.prod(s))
        return buf[self.off[name]: self.off[name] + n].view(s)

    def p(self, name):
        return self._v(self.P, name) if name in self.off else None

    def g(self, name):
        return self._v(self.G, name) if name in self.off else None

    def w16(self, name):
        return self._v(self.P16, name)

    def wT16(self, name):
        s = self.shape[name]
        return self.PT16[self.off[name]: self.off[name] + s[0] * s[1]].view(s[1], s[0])

    def refresh_operands(self):
        """bf16 (and transposed) copies of the parameters - after construction, load_state_dict or any edit of the fp32 values."""
        self.P16.copy_(self.P)
        self.ops.weights_transpose(self.P, self.PT16, self._jobs, self._n_jobs, self._n_tiles)
        self.module.model.refresh()

    # ---- concurrency: independent chains on side streams (branches of the graph under capture) ------------------------------
    def _branches(self, n):
        eng = self

        class _Br:
            def __enter__(self):
                self.on = eng.device.type == "cuda" and eng.concurrent
                if self.on:
                    self.cur = torch.cuda.current_stream(eng.device)
                    while len(eng._streams) < n:
                        eng._streams.append(torch.cuda.Stream(eng.device))
                    for st in eng._streams[:n]:
                        st.wait_stream(self.cur)
                return self

            def __call__(self, i):
                import contextlib
                return torch.cuda.stream(eng._streams[i]) if self.on else contextlib.nullcontext()

            def __exit__(self, *exc):
                if self.on:
                    for st in eng._streams[:n]:
                        self.cur.wait_stream(st)
                return False
        return _Br()

    def _wgrad_side(self, fn, *tensors):
        """Weight-gradient products feed nothing downstream: they run on a side stream next to the data-gradient chain and are
        joined once at the end of the backward pass.  Their operands are kept alive until then."""
        if self.device.type != "cuda" or not self.concurrent:
            return fn()
        cur = torch.cuda.current_stream(self.device)
        key = cur.cuda_stream
        st = self._wstreams.get(key)
        if st is None:
            st = self._wstreams[key] = torch.cuda.Stream(self.device)
        self._wused.add(key)
        self._keep.extend(tensors)
        st.wait_stream(cur)
        with torch.cuda.stream(st):
            fn()

    def _join_wgrad(self):
        if self._wused:
            cur = torch.cuda.current_stream(self.device)
            for key in self._wused:
                cur.wait_stream(self._wstreams[key])
            self._wused.clear()
        self._keep.clear()

    # ---- nn.Linear on packed rows ----------------------------------------------------------------------------------------
    def _lin_fwd(self, x16, name, out, mode, aux=None, resid=None, tadd=None, row_jet=None):
        self.ops.gemm(x16, self.w16(name + ".weight"), out, self.p(name + ".bias"), mode, aux=aux, resid=resid, tadd=tadd, row_jet=row_jet)

    def _ksplit(self, plan, n_out, n_in):
        tiles = ((n_out + 127) // 128) * ((n_in + 127) // 128)
        kb = (plan.M + 63) // 64
        return max(1, min(kb // 4, max(1, 296 // tiles)))

    def _lin_bwd(self, plan, dy16, x16, name, dx=None, dx_mode=1, dy_src=None, bias_done=False, aux=None):
        """dy16 [M, N], x16 [M, K] bf16 row-major (column slices allowed).  dW += dy^T x (both operands read MN-major by the
        tensor cores: no transposed copies), db += colsum(dy) unless the producer of dy16 already added it (bias_done), dx = dy W
        (dx_mode 4: times GELU'(aux), the data gradient straight through the activation).  dy_src: fp32 source first cast into dy16."""
        ops = self.ops
        N, K = self.shape[name + ".weight"]
        db = None if bias_done else self.g(name + ".bias")
        if dy_src is not None:
            ops.cast_transpose(dy_src, out=dy16, colsum=db)
            db = None
        ks = self._ksplit(plan, N, K)

        def wgrad():
            if db is not None:
                ops.cast_transpose(dy16, colsum=db)
            ops.gemm_tn(dy16, x16, self.g(name + ".weight"), ks)
        self._wgrad_side(wgrad, dy16, x16)
        if dx is not None:
            ops.gemm(dy16, self.wT16(name + ".weight"), dx, None, dx_mode, aux=aux)

    # ---- one SelfAttnBlock (reference networks/attention.py:23-26) on columns of the 256-wide residual buffers ------------
    def _block_fwd(self, plan, Rin, R1, R2, pre, tadd):
        ops, dev, M = self.ops, self.device, plan.M
        C = Rin.shape[1]
        H, hs, I = self.H, C // self.H, self.I
        bf = dict(device=dev, dtype=torch.bfloat16)
        s = {"Rin": Rin, "R1": R1, "C": C, "pre": pre}
        s["a1"], s["m1"], s["r1"] = torch.empty(M, C, **bf), torch.empty(M, device=dev), torch.empty(M, device=dev)
        ops.ln_fwd(Rin, self.p(pre + ".ln1.weight"), self.p(pre + ".ln1.bias"), s["m1"], s["r1"], out16=s["a1"])
        s["qkv"] = torch.empty(M, 3 * C, **bf)
        if self.cfg.qk_layernorm:                              # c_attn with the q / k LayerNorm in its epilogue
            qkn = torch.empty(M, 2 * C, **bf)
            ops.gemm_qkv(s["a1"], self.w16(pre + ".attn.c_attn.weight"), self.p(pre + ".attn.c_attn.bias"), s["qkv"], qkn, H,
                         self.p(pre + ".attn.q_layernorm.weight"), self.p(pre + ".attn.q_layernorm.bias"),
                         self.p(pre + ".attn.k_layernorm.weight"), self.p(pre + ".attn.k_layernorm.bias"))
            s["qn"], s["kn"] = qkn[:, :C], qkn[:, C:]
        else:
            self._lin_fwd(s["a1"], pre + ".attn.c_attn", s["qkv"], 0)
            s["qn"], s["kn"] = s["qkv"][:, :C], s["qkv"][:, C:2 * C]
        s["o"] = torch.zeros(M, C, **bf) if plan.padded else torch.empty(M, C, **bf)      # rows of no jet are never written
        s["stats"] = torch.empty(M, H, 2, device=dev)
        ops.attn_tc_fwd(s["qn"], s["kn"], s["qkv"][:, 2 * C:], hs, plan.items, plan.n_items, plan.grid_items, plan.row_jet, plan.jet_off,
                        s["stats"], s["o"])
        if plan.has_big:                                      # jets of 129 ... 150 particles: CUDA-core kernels, probabilities kept
            s["P"] = torch.empty(max(plan.sum_n2 * H, 1), **bf)
            ops.attn_fwd(s["qn"], s["kn"], s["qkv"][:, 2 * C:], plan.jet_off, plan.p_off, plan.B, H, hs, plan.nmax, s["o"], s["P"], min_n=128)
        self._lin_fwd(s["o"], pre + ".attn.c_proj", R1, 5, resid=Rin)           # x = x + attn(ln1(x)), written out of place
        s["a2"], s["m2"], s["r2"] = torch.empty(M, C, **bf), torch.empty(M, device=dev), torch.empty(M, device=dev)
        ops.ln_fwd(R1, self.p(pre + ".ln2.weight"), self.p(pre + ".ln2.bias"), s["m2"], s["r2"], out16=s["a2"])
        s["z"], s["hh"] = torch.empty(M, I, **bf), torch.empty(M, I, **bf)
        self._lin_fwd(s["a2"], pre + ".ffw.c_fc", s["z"], 3, aux=s["hh"])      # z and GELU(z) leave the same epilogue
        self._lin_fwd(s["hh"], pre + ".ffw.c_proj", R2, 5, resid=R1, tadd=tadd, row_jet=plan.row_jet)   # x = x + ffw(ln2(x)) + time embedding
        return s

    def _block_bwd(self, plan, s, G, g16=None, nxt=None):
        """G [M, C] fp32: gradient w.r.t. the block's output on entry, w.r.t. its input on exit.  g16: its bf16 copy when the
        producer of G already made one (and added its column sums to ffw.c_proj.bias); nxt = (bf16 buffer, bias gradient) that
        the last LayerNorm backward of this block fills for the block processed next."""
        ops, dev, M = self.ops, self.device, plan.M
        C, pre, H, I = s["C"], s["pre"], self.H, self.I
        hs = C // H
        bf = dict(device=dev, dtype=torch.bfloat16)
        dz = torch.empty(M, I, **bf)
        if g16 is None:
            self._lin_bwd(plan, torch.empty(M, C, **bf), s["hh"], pre + ".ffw.c_proj", dx=dz, dx_mode=4, dy_src=G, aux=s["z"])
        else:
            self._lin_bwd(plan, g16, s["hh"], pre + ".ffw.c_proj", dx=dz, dx_mode=4, bias_done=True, aux=s["z"])
        da = torch.empty(M, C, device=dev)
        self._lin_bwd(plan, dz, s["a2"], pre + ".ffw.c_fc", dx=da, dx_mode=1)
        do, G16b = torch.empty(M, C, **bf), torch.empty(M, C, **bf)
        ops.ln_bwd(da, s["R1"], s["m2"], s["r2"], self.p(pre + ".ln2.weight"), G, self.g(pre + ".ln2.weight"), self.g(pre + ".ln2.bias"),
                   accumulate=True, dx16=G16b, dxsum=self.g(pre + ".attn.c_proj.bias"))
        self._lin_bwd(plan, G16b, s["o"], pre + ".attn.c_proj", dx=do, dx_mode=0, bias_done=True)
        dqkv = torch.zeros(M, 3 * C, **bf) if plan.padded else torch.empty(M, 3 * C, **bf)
        ops.attn_tc_bwd(do, s["qn"], s["kn"], s["qkv"][:, 2 * C:], hs, plan.items, plan.n_items, plan.grid_items, plan.row_jet, plan.jet_off,
                        s["stats"], dqkv)
        if plan.has_big:
            ops.attn_bwd(do, s["o"], s["P"], s["qn"], s["kn"], s["qkv"][:, 2 * C:], plan.jet_off, plan.p_off, plan.B, H, hs, plan.nmax, dqkv, C,
                         min_n=128)
        if self.cfg.qk_layernorm:
            ops.qkln_bwd(dqkv, s["qkv"], C, H, self.p(pre + ".attn.q_layernorm.weight"), self.p(pre + ".attn.k_layernorm.weight"),
                         self.g(pre + ".attn.q_layernorm.weight"), self.g(pre + ".attn.q_layernorm.bias"),
                         self.g(pre + ".attn.k_layernorm.weight"), self.g(pre + ".attn.k_layernorm.bias"))
        self._lin_bwd(plan, dqkv, s["a1"], pre + ".attn.c_attn", dx=da, dx_mode=1)
        ops.ln_bwd(da, s["Rin"], s["m1"], s["r1"], self.p(pre + ".ln1.weight"), G, self.g(pre + ".ln1.weight"), self.g(pre + ".ln1.bias"),
                   accumulate=True, dx16=None if nxt is None else nxt[0], dxsum=None if nxt is None else nxt[1])

    def _chain_bwd(self, plan, chain, G, g16, before=None):
        """Backward through a chain of blocks (last block first); g16 = bf16 copy of G for the first one processed."""
        order = list(reversed(chain))
        for i, s in enumerate(order):
            if before is not None:
                before()
            nxt = None
            if i + 1 < len(order):
                nxt = (torch.empty(plan.M, s["C"], device=self.device, dtype=torch.bfloat16), self.g(order[i + 1]["pre"] + ".ffw.c_proj.bias"))
            self._block_bwd(plan, s, G, g16, nxt)
            g16 = None if nxt is None else nxt[0]

    # ---- encoder forward on packed rows (reference ParticleTransformers.py:62-122 / 177-210) ------------------------------
    def _forward(self, plan, xs, ks, t):
        ops, dev, M, B = self.ops, self.device, plan.M, plan.B
        E, h, I = self.E, self.h, self.I
        T = "model.transformer."
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = lambda *s: torch.empty(*s, device=dev)
        c = {}
        # time embeddings: ParticleFormer adds the same 128-wide row to both streams, the fused encoder one 256-wide row
        temb = f32(B, 256)
        ops.time_embed(t, h if self.pf else E, self.pf, temb, perm=plan.perm)
        c["temb"] = temb
        if self.pf:
            c["temb2"] = f32(B, E)
            w = self.p(T + "time_expand.weight")
            ops.sgemm(temb, 256, 1, w, 1, h, c["temb2"], B, E, h, bias=self.p(T + "time_expand.bias"))
        # embeddings
        c["h0"], c["g0"] = torch.empty(M, E, **bf), torch.empty(M, E, **bf)
        ops.embed_x_fwd(xs, self.p(T + "wxe.0.weight"), self.p(T + "wxe.0.bias"), c["h0"])
        ops.embed_y_fwd(ks, self.p(T + "wye.0.weight"), c["g0"])
        c["u"] = f32(M, 256)                                      # pre-LayerNorm embeddings x | y
        self._lin_fwd(c["h0"], T + "wxe.2", c["u"][:, :h], 1)
        self._lin_fwd(c["g0"], T + "wye.2", c["u"][:, h:], 1)
        R = f32(M, 256)
        c["em"], c["er"] = f32(2, M), f32(2, M)
        for gi, nm in enumerate(("ln1_x", "ln1_y")):
            cols = slice(gi * h, (gi + 1) * h)
            ops.ln_fwd(c["u"][:, cols], self.p(T + nm + ".weight"), self.p(T + nm + ".bias"), c["em"][gi], c["er"][gi],
                       tadd=temb[:, cols], row_jet=plan.row_jet, out32=R[:, cols])
        c["skip"] = R
        blocks: List[dict] = []
        if self.pf:
            # the continuous and the discrete stream do not meet before ln2_x / ln2_y: two independent chains of blocks, run
            # concurrently (two branches of the CUDA graph) - each one alone leaves most of the chip idle
            Rs = [R] + [(f32(M, 256), f32(M, 256)) for _ in range(self.cfg.n_layer)]
            chains = [[], []]
            with self._branches(2) as br:
                for gi, nm in enumerate(("blocks_x", "blocks_y")):
                    cols = slice(gi * h, (gi + 1) * h)
                    with br(gi):
                        Rin = R
                        for i in range(self.cfg.n_layer):
                            R1, R2 = Rs[i + 1]
                            chains[gi].append(self._block_fwd(plan, Rin[:, cols], R1[:, cols], R2[:, cols], f"{T}{nm}.{i}", temb[:, cols]))
                            Rin = R2
            blocks = chains[0] + chains[1]
            R = Rs[-1][1]
            # x = ln2_x(x + x_skip) | y = ln2_y(y + y_skip); z = cat + time_expand(temb)
            c["Rmid"] = R
            Z = f32(M, 256)
            c["mm"], c["mr"] = f32(2, M), f32(2, M)
            for gi, nm in enumerate(("ln2_x", "ln2_y")):
                cols = slice(gi * h, (gi + 1) * h)
                ops.ln_fwd(R[:, cols], self.p(T + nm + ".weight"), self.p(T + nm + ".bias"), c["mm"][gi], c["mr"][gi], add=c["skip"][:, cols],
                           tadd=c["temb2"][:, cols], row_jet=plan.row_jet, out32=Z[:, cols])
            R = Z
            main, main_t, n_main = "blocks_fuse", c["temb2"], self.cfg.n_layer_fused
        else:
            main, main_t, n_main = "blocks", temb, self.cfg.n_layer
        c["n_stream_blocks"] = len(blocks)
        for i in range(n_main):
            R1, R2 = f32(M, 256), f32(M, 256)
            blocks.append(self._block_fwd(plan, R, R1, R2, f"{T}{main}.{i}", main_t))
            R = R2
        c["blocks"], c["Rlast"] = blocks, R
        # final LayerNorm(s) over (z + skip), heads
        c["xf"] = torch.empty(M, 256, **bf)
        if self.pf:
            c["fm"], c["fr"] = f32(2, M), f32(2, M)
            for gi, nm in enumerate(("ln3_x", "ln3_y")):
                cols = slice(gi * h, (gi + 1) * h)
                ops.ln_fwd(R[:, cols], self.p(T + nm + ".weight"), self.p(T + nm + ".bias"), c["fm"][gi], c["fr"][gi], add=c["skip"][:, cols],
                           out16=c["xf"][:, cols])
        else:
            c["fm"], c["fr"] = f32(1, M), f32(1, M)
            ops.ln_fwd(R, self.p(T + "ln2.weight"), self.p(T + "ln2.bias"), c["fm"][0], c["fr"][0], add=c["skip"], out16=c["xf"])
        c["zh"], c["hh"] = torch.empty(M, 2 * I, **bf), torch.empty(M, 2 * I, **bf)
        for gi, nm in enumerate(("head_x.0", "head_y.0")):
            self._lin_fwd(c["xf"][:, gi * h:(gi + 1) * h], T + nm, c["zh"][:, gi * I:(gi + 1) * I], 0)
        ops.gelu_fwd(c["zh"], c["hh"])
        c["vt"], c["logits"] = f32(M, 3), f32(M, self.V)
        ops.head_fwd(c["hh"], I, self.p(T + "head_x.2.weight"), self.p(T + "head_x.2.bias"), self.p(T + "head_y.2.weight"),
                     self.p(T + "head_y.2.bias"), c["vt"], c["logits"])
        return c

    def _loss(self, plan, c, tgt, k1p, t, want_grads):
        """MultiTaskLoss (reference model/MMF.py:152-168, 203-233) and, for the backward pass, d loss / d (vt, logits, u)."""
        ops, dev, B, E = self.ops, self.device, plan.B, self.E
        f32 = lambda *s: torch.empty(*s, device=dev)
        l1, l2 = f32(B), f32(B)
        ops.loss_fwd(c["vt"], c["logits"], tgt, k1p, plan.jet_off, B, self.V, l1, l2)
        u = None
        if self.cfg.multitask_loss == "weighted":             # u[b, :] = loss_weights for every jet (reference model/MMF.py:219-223)
            if self._ones is None or self._ones.shape[0] < B:
                self._ones = torch.ones(max(B, 256), device=dev)
            u = f32(B, 2)
            ops.sgemm(self._ones, 1, 0, self.p("loss_combine.loss_weights"), 0, 1, u, B, 2, 1)
        if self.cfg.multitask_loss == "time-weighted":
            N = "loss_combine.uncertainty_net."
            c["ue"], c["ua"], c["uh"], u = f32(B, E), f32(B, E), f32(B, E), f32(B, 2)
            ops.time_embed(t, E, False, c["ue"], perm=plan.perm)
            ops.sgemm(c["ue"], E, 1, self.p(N + "c_fc.weight"), 1, E, c["ua"], B, E, E, bias=self.p(N + "c_fc.bias"))
            ops.gelu_fwd(c["ua"], c["uh"])
            ops.sgemm(c["uh"], E, 1, self.p(N + "c_proj.weight"), 1, E, u, B, 2, E, bias=self.p(N + "c_proj.bias"))
        out5, gl1, gl2 = f32(5), f32(B), f32(B)
        c["du"] = f32(B, 2) if u is not None else None
        ops.loss_combine(l1, l2, u, out5, gl1, gl2, c["du"])
        if want_grads:
            c["dvt"], c["dlog"] = f32(plan.M, 3), f32(plan.M, self.V)
            ops.loss_bwd(c["vt"], c["logits"], tgt, k1p, plan.row_jet, plan.jet_off, gl1, gl2, self.V, c["dvt"], c["dlog"])
        return out5

    # ---- backward ---------------------------------------------------------------------------------------------------------
    def _backward(self, plan, c, xs, ks):
        ops, dev, M, B = self.ops, self.device, plan.M, plan.B
        E, h, I = self.E, self.h, self.I
        T = "model.transformer."
        bf = dict(device=dev, dtype=torch.bfloat16)
        f32 = lambda *s: torch.empty(*s, device=dev)
        if c["du"] is not None and self.cfg.multitask_loss == "weighted":
            ops.cast_transpose(c["du"], colsum=self.g("loss_combine.loss_weights"))
        # uncertainty net
        if c["du"] is not None and self.cfg.multitask_loss == "time-weighted":
            N = "loss_combine.uncertainty_net."
            du = c["du"]
            ops.sgemm(du, 1, 2, c["uh"], E, 1, self.g(N + "c_proj.weight"), 2, E, B, accumulate=True)
            ops.cast_transpose(du, colsum=self.g(N + "c_proj.bias"))
            dh = f32(B, E)
            ops.sgemm(du, 2, 1, self.p(N + "c_proj.weight"), E, 1, dh, B, E, 2)
            ops.gelu_bwd(dh, c["ua"], dh)
            ops.sgemm(dh, 1, E, c["ue"], E, 1, self.g(N + "c_fc.weight"), E, E, B, accumulate=True)
            ops.cast_transpose(dh, colsum=self.g(N + "c_fc.bias"))
        # heads
        dzh = torch.empty(M, 2 * I, **bf)
        ops.head_bwd(c["dvt"], c["dlog"], c["hh"], c["zh"], I, self.p(T + "head_x.2.weight"), self.p(T + "head_y.2.weight"), dzh,
                     self.g(T + "head_x.2.weight"), self.g(T + "head_x.2.bias"), self.g(T + "head_y.2.weight"), self.g(T + "head_y.2.bias"))
        dxf = f32(M, 256)
        for gi, nm in enumerate(("head_x.0", "head_y.0")):
            self._lin_bwd(plan, dzh[:, gi * I:(gi + 1) * I], c["xf"][:, gi * h:(gi + 1) * h], T + nm, dx=dxf[:, gi * h:(gi + 1) * h], dx_mode=1)
        # final LayerNorm(s): d/d(z + skip) goes to the residual stream and to the skip connection alike; the same kernels hand the
        # first block of the backward chain its bf16 operand and bias column sums
        G, R = f32(M, 256), c["Rlast"]
        blocks, ns = c["blocks"], c["n_stream_blocks"]
        main = blocks[ns:]
        g16 = torch.empty(M, 256, **bf)
        last_bias = self.g(main[-1]["pre"] + ".ffw.c_proj.bias")
        if self.pf:
            for gi, nm in enumerate(("ln3_x", "ln3_y")):
                cols = slice(gi * h, (gi + 1) * h)
                ops.ln_bwd(dxf[:, cols], R[:, cols], c["fm"][gi], c["fr"][gi], self.p(T + nm + ".weight"), G[:, cols], self.g(T + nm + ".weight"),
                           self.g(T + nm + ".bias"), add=c["skip"][:, cols], dx16=g16[:, cols], dxsum=None if last_bias is None else last_bias[cols])
        else:
            ops.ln_bwd(dxf, R, c["fm"][0], c["fr"][0], self.p(T + "ln2.weight"), G, self.g(T + "ln2.weight"), self.g(T + "ln2.bias"), add=c["skip"],
                       dx16=g16, dxsum=last_bias)
        Gskip = f32(M, 256)
        ops.add(Gskip, G)
        if self.pf:
            dt2 = torch.zeros(B, E, device=dev)
            self._chain_bwd(plan, main, G, g16, before=lambda: ops.jet_sum(G, plan.jet_off, B, dt2, accumulate=True))
            ops.jet_sum(G, plan.jet_off, B, dt2, accumulate=True)
            w = T + "time_expand."
            ops.sgemm(dt2, 1, E, c["temb"], 256, 1, self.g(w + "weight"), E, h, B, accumulate=True)
            ops.cast_transpose(dt2, colsum=self.g(w + "bias"))
            G2 = f32(M, 256)
            chains = [blocks[: ns // 2], blocks[ns // 2: ns]]
            g16s = [torch.empty(M, h, **bf), torch.empty(M, h, **bf)]
            for gi, nm in enumerate(("ln2_x", "ln2_y")):
                cols = slice(gi * h, (gi + 1) * h)
                ops.ln_bwd(G[:, cols], c["Rmid"][:, cols], c["mm"][gi], c["mr"][gi], self.p(T + nm + ".weight"), G2[:, cols],
                           self.g(T + nm + ".weight"), self.g(T + nm + ".bias"), add=c["skip"][:, cols], dx16=g16s[gi],
                           dxsum=self.g(chains[gi][-1]["pre"] + ".ffw.c_proj.bias"))
            ops.add(Gskip, Gskip, G2)
            G = G2
            with self._branches(2) as br:
                for gi in range(2):
                    with br(gi):
                        self._chain_bwd(plan, chains[gi], G[:, gi * h:(gi + 1) * h], g16s[gi])
        else:
            self._chain_bwd(plan, main, G, g16)
        ops.add(G, G, Gskip)
        # embeddings
        du = f32(M, 256)
        for gi, nm in enumerate(("ln1_x", "ln1_y")):
            cols = slice(gi * h, (gi + 1) * h)
            ops.ln_bwd(G[:, cols], c["u"][:, cols], c["em"][gi], c["er"][gi], self.p(T + nm + ".weight"), du[:, cols], self.g(T + nm + ".weight"),
                       self.g(T + nm + ".bias"))
        du16 = torch.empty(M, 256, **bf)
        dh0, dg0 = torch.empty(M, E, **bf), torch.empty(M, E, **bf)
        self._lin_bwd(plan, du16[:, :h], c["h0"], T + "wxe.2", dx=dh0, dx_mode=0, dy_src=du[:, :h])
        self._lin_bwd(plan, du16[:, h:], c["g0"], T + "wye.2", dx=dg0, dx_mode=0, dy_src=du[:, h:])
        ops.embed_x_bwd(dh0, xs, self.p(T + "wxe.0.weight"), self.p(T + "wxe.0.bias"), self.g(T + "wxe.0.weight"), self.g(T + "wxe.0.bias"))
        ops.embed_y_bwd(dg0, ks, self.p(T + "wye.0.weight"), self.g(T + "wye.0.weight"))
        self._join_wgrad()

    # ---- public API ---------------------------------------------------------------------------------------------------------
    def _prepare(self, batch: DataCoupling, time, z, u):
        cfg, dev, mod = self.cfg, self.device, self.module
        B, V, eps = len(batch), cfg.vocab_size, cfg.time_eps
        if time is None:
            time = eps + (1.0 - eps) * torch.rand(B, device=dev)
        time = time.to(dev, torch.float32).contiguous()
        plan = _Plan(batch.target.mask, dev, upload=not self.use_graphs)
        # host batches (a DataLoader's, ideally pinned) are copied without blocking; the plan above came from the host mask
        up = lambda tm: tm._map(lambda x: x.to(dev, non_blocking=True))
        tgt, src = up(batch.target), up(batch.source) if batch.source is not None else TensorMultiModal()
        if not src.has_continuous:                       # reference model/CFM.py:175-177
            src.continuous = torch.randn_like(tgt.continuous) * tgt.mask
        if not src.has_discrete:                         # reference model/MJB.py:201-203
            src.discrete = torch.randint_like(tgt.discrete, 1, V) * tgt.mask
        xt, kt = _abi.bridge_sample(src.continuous, tgt.continuous, src.discrete, tgt.discrete, time, cfg.sigma, cfg.beta, V,
                                    z=None if z is None else z.to(dev), u=None if u is None else u.to(dev), seed=mod.seed,
                                    first_global_jet=mod._jet_cursor)
        mod._jet_cursor += B
        self.last_plan = plan
        if self.use_graphs:
            rows = max(self.graph_rows, (plan.M + self.graph_rows - 1) // self.graph_rows * self.graph_rows)
            key = (plan.B, plan.D, rows, plan.has_big)
            slot = self._slots.pop(key, None)
            if slot is None:
                while len(self._slots) >= self.max_graphs:      # every graph owns ~1 GB of activations: keep the most recent few
                    self._slots.pop(next(iter(self._slots)))
                slot = _GraphSlot(plan.B, plan.D, rows, plan.has_big, dev)
            self._slots[key] = slot                             # (re-inserted last: the dict is the LRU order)
            slot.load(plan)
            slot.t.copy_(time)
            xs, ks, tg, k1p, run = slot.xs, slot.ks, slot.tg, slot.k1p, slot
        else:
            M = plan.M
            xs, tg = torch.empty(M, 3, device=dev), torch.empty(M, 3, device=dev)
            ks, k1p = torch.empty(M, device=dev, dtype=torch.int32), torch.empty(M, device=dev, dtype=torch.int32)
            run = plan
        self.ops.pack(xt.contiguous(), kt.contiguous(), src.continuous.contiguous(), tgt.continuous.contiguous(), tgt.discrete.contiguous(),
                      plan.row_slot, V, xs, ks, tg, k1p, self._err)
        return run, time, xs, ks, tg, k1p

    def _fwd_bwd(self, plan, xs, ks, tg, k1p, t):
        self.G.zero_()
        c = self._forward(plan, xs, ks, t)
        out5 = self._loss(plan, c, tg, k1p, t, True)
        self._backward(plan, c, xs, ks)
        return out5

    def loss_and_grad(self, batch: DataCoupling, time=None, z=None, u=None):
        """One forward + backward pass.  Returns the reference's (loss, loss_mse, loss_ce, w_mse, w_ce) as a 5-vector on the
        device; the gradients are in ``self.G`` (= every parameter's ``.grad``).  With ``use_graphs`` the ~600 kernel launches
        of the pass are ONE CUDA graph per (jets, padded row count), captured on first use and replayed afterwards."""
        plan, t, xs, ks, tg, k1p = self._prepare(batch, time, z, u)
        if not self.use_graphs:
            return self._fwd_bwd(plan, xs, ks, tg, k1p, t)
        slot = plan
        if slot.graph is None:
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                 # eager warm-up off the capture: one-time attribute opt-ins, allocator
                self._fwd_bwd(slot, xs, ks, tg, k1p, slot.t)
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                slot.out5 = self._fwd_bwd(slot, xs, ks, tg, k1p, slot.t)
            slot.graph = graph
        slot.graph.replay()
        self.ops.launches += 1
        return slot.out5

    def loss_only(self, batch: DataCoupling, time=None, z=None, u=None):
        graphs, self.use_graphs = self.use_graphs, False
        try:
            plan, t, xs, ks, tg, k1p = self._prepare(batch, time, z, u)
        finally:
            self.use_graphs = graphs
        c = self._forward(plan, xs, ks, t)
        return self._loss(plan, c, tg, k1p, t, False)

    def optimizer_step(self, lr: Optional[float] = None):
        """DDP average (one all-reduce of the flat gradient), gradient-norm clipping, Adam, refreshed bf16 operands."""
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
            if world > 1:
                torch.distributed.all_reduce(self.G)
        self.step_count += 1
        ops = self.ops
        sumsq = None
        if self.max_norm > 0:
            ops.sumsq(self.G, self._sumsq)
            sumsq = self._sumsq
        ops.adam(self.P, self.G, self.m, self.v, self.lr if lr is None else lr, self.betas[0], self.betas[1], self.eps, self.step_count,
                 sumsq=sumsq, max_norm=self.max_norm, grad_scale=1.0 / world, p16=self.P16)
        ops.weights_transpose(self.P, self.PT16, self._jobs, self._n_jobs, self._n_tiles)
        self.module.model.refresh()

    def train_step(self, batch: DataCoupling, lr: Optional[float] = None, time=None, z=None, u=None):
        out5 = self.loss_and_grad(batch, time, z, u)
        self.optimizer_step(lr)
        return out5

    def grad_norm(self) -> float:
        self.ops.sumsq(self.G, self._sumsq)
        return float(self._sumsq[0].sqrt())

    def check_tokens(self):
        if int(self._err.item()):
            self._err.zero_()
            raise RuntimeError("Values in `k` outside of bound [0, vocab_size)")      # reference model/MJB.py:177-182


def lr_schedule(cfg, epochs: int) -> List[float]:
    """Learning rate of every epoch as the reference's SequentialLR(LinearLR warm-up, CosineAnnealingLR) yields it
    (model/MMF.py:79-110) - produced by the same torch schedulers on a dummy parameter."""
    from torch.optim.lr_scheduler import CosineAnnealingLR, LinearLR, SequentialLR
    opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=cfg.lr)
    cosine = CosineAnnealingLR(opt, T_max=max(cfg.max_epochs - cfg.warmup_epochs, 1), eta_min=cfg.lr_final, last_epoch=-1)
    warm = LinearLR(opt, start_factor=0.01, end_factor=1.0, total_iters=cfg.warmup_epochs)
    sched = SequentialLR(opt, schedulers=[warm, cosine], milestones=[cfg.warmup_epochs])
    out = []
    for _ in range(epochs):
        out.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    return out
