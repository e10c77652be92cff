"""Test seam for the HOST logic of the training step: a plain-torch stand-in with the interface of
``mmf_b200._train_abi.Ops`` (same argument order, same in-place / accumulate conventions, bf16 storage where the kernels store
bf16).  ``tests/test_training_host_logic.py`` injects it into ``TrainEngine`` so that the sequencing of the forward and
backward programs (which operator on which buffer, in which order) is checked against torch autograd on the CPU, without a GPU
and without the library.  It is test infrastructure only - nothing under multimodal-flows_b200/ imports it."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def _f(t):
    return t.float()


class MockOps:
    def __init__(self):
        self.launches = 0

    def gemm(self, A, B, C, bias=None, mode=0, ksplit=1, aux=None, resid=None, tadd=None, row_jet=None):
        r = _f(A) @ _f(B).T
        if bias is not None:
            r = r + bias
        if mode == 2:
            C += r
        elif mode == 3:
            C.copy_(r.to(C.dtype))
            aux.copy_(F.gelu(_f(C)).to(aux.dtype))
        elif mode == 4:
            zz = _f(aux).clone().requires_grad_(True)
            C.copy_(torch.autograd.grad(F.gelu(zz), zz, r)[0].to(C.dtype))
        elif mode == 5:
            r = r + resid
            if tadd is not None:
                r = r + tadd[row_jet.long()]
            C.copy_(r)
        else:
            C.copy_(r.to(C.dtype))

    def gemm_qkv(self, A, W, bias, qkv, qkn, H, qg, qb, kg, kb):
        self.gemm(A, W, qkv, bias, 0)
        C = W.shape[0] // 3
        self.qkln_fwd(qkv, C, H, qg, qb, kg, kb, qkn[:, :C], qkn[:, C:])

    def gemm_tn(self, A, B, C, ksplit=1):
        C += _f(A).T @ _f(B)

    def sgemm(self, A, sam, sak, B, sbk, sbn, C, M, N, K, bias=None, accumulate=False):
        a = torch.as_strided(A, (M, K), (sam, sak))
        b = torch.as_strided(B, (K, N), (sbk, sbn))
        r = a @ b
        if bias is not None:
            r = r + bias
        c = C.view(M, N) if C.dim() == 1 else C
        if accumulate:
            c += r
        else:
            c.copy_(r)

    def cast_transpose(self, x, out=None, outT=None, colsum=None):
        rows = x.shape[0]
        if out is not None:
            out.copy_(x.to(torch.bfloat16))
        if outT is not None:
            outT.zero_()
            outT[:, :rows] = x.to(torch.bfloat16).T
        if colsum is not None:
            colsum += _f(x).sum(0)

    def weights_transpose(self, params, paramsT, jobs, n_jobs, n_tiles):
        rec = jobs.cpu().numpy().view([("src", "<i8"), ("dst", "<i8"), ("rows", "<i4"), ("cols", "<i4"), ("tile0", "<i4"), ("pad", "<i4")])
        for j in rec:
            n = int(j["rows"]) * int(j["cols"])
            src = params[int(j["src"]): int(j["src"]) + n].view(int(j["rows"]), int(j["cols"]))
            paramsT[int(j["dst"]): int(j["dst"]) + n] = src.T.contiguous().to(torch.bfloat16).flatten()

    def time_embed(self, t, dim, dup, out, perm=None):
        if perm is not None:
            t = t[perm.long()]
        half = dim // 2
        f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
        e = t[:, None] * f[None]
        e = torch.cat([e.sin(), e.cos()], 1)
        out[:, :dim] = e
        if dup:
            out[:, dim:2 * dim] = e

    def embed_x_fwd(self, xs, w0, b0, h):
        h.copy_(F.gelu(xs @ w0.T + b0).to(h.dtype))

    def embed_x_bwd(self, dh, xs, w0, b0, dw0, db0):
        z = (xs @ w0.T + b0).requires_grad_(True)
        dz = torch.autograd.grad(F.gelu(z), z, _f(dh))[0]
        dw0 += dz.T @ xs
        db0 += dz.sum(0)

    def embed_y_fwd(self, ks, emb, g):
        g.copy_(F.gelu(emb[ks.long()]).to(g.dtype))

    def embed_y_bwd(self, dg, ks, emb, demb):
        e = emb[ks.long()].clone().requires_grad_(True)
        de = torch.autograd.grad(F.gelu(e), e, _f(dg))[0]
        demb.index_add_(0, ks.long(), de)

    def ln_fwd(self, x, g, b, mean, rstd, add=None, tadd=None, row_jet=None, out16=None, out32=None):
        v = x + add if add is not None else x
        mu = v.mean(1)
        rs = torch.rsqrt(v.var(1, unbiased=False) + 1e-5)
        mean.copy_(mu)
        rstd.copy_(rs)
        y = (v - mu[:, None]) * rs[:, None] * g + (b if b is not None else 0)
        if tadd is not None:
            y = y + tadd[row_jet.long()]
        if out32 is not None:
            out32.copy_(y)
        if out16 is not None:
            out16.copy_(y.to(torch.bfloat16))

    def ln_bwd(self, dy, x, mean, rstd, g, dx, dg, db, add=None, accumulate=False, dx16=None, dxsum=None):
        v = x + add if add is not None else x
        xh = (v - mean[:, None]) * rstd[:, None]
        dxh = dy * g
        d = rstd[:, None] * (dxh - dxh.mean(1, keepdim=True) - xh * (dxh * xh).mean(1, keepdim=True))
        if accumulate:
            dx += d
        else:
            dx.copy_(d)
        dg += (dy * xh).sum(0)
        if db is not None:
            db += dy.sum(0)
        if dx16 is not None:
            dx16.copy_(dx.to(torch.bfloat16))
        if dxsum is not None:
            dxsum += dx.sum(0)

    def qkln_fwd(self, qkv, C, H, qg, qb, kg, kb, qn, kn):
        M, hs = qkv.shape[0], C // H
        q = F.layer_norm(_f(qkv[:, :C]).view(M, H, hs), (hs,), qg, qb, 1e-5).reshape(M, C)
        k = F.layer_norm(_f(qkv[:, C:2 * C]).view(M, H, hs), (hs,), kg, kb, 1e-5).reshape(M, C)
        qn.copy_(q.to(qn.dtype))
        kn.copy_(k.to(kn.dtype))

    def qkln_bwd(self, dqkv, qkv, C, H, qg, kg, dqg, dqb, dkg, dkb):
        M, hs = qkv.shape[0], C // H
        for w, (gam, dgam, dbet) in enumerate(((qg, dqg, dqb), (kg, dkg, dkb))):
            x = _f(qkv[:, w * C:(w + 1) * C]).view(M, H, hs).clone().requires_grad_(True)
            gp = gam.clone().requires_grad_(True)
            bp = torch.zeros(hs, requires_grad=True)
            y = F.layer_norm(x, (hs,), gp, bp, 1e-5)
            gx, gg, gb = torch.autograd.grad(y, (x, gp, bp), _f(dqkv[:, w * C:(w + 1) * C]).view(M, H, hs))
            dqkv[:, w * C:(w + 1) * C] = gx.reshape(M, C).to(dqkv.dtype)
            dgam += gg
            if dbet is not None:
                dbet += gb

    def _heads(self, t, r, n, H, hs):
        return _f(t[r]).view(n, H, hs).transpose(0, 1)

    def attn_tc_fwd(self, qn, kn, v, hs, items, n_items, grid_items, row_jet, jet_off, stats, o):
        H, c = qn.shape[1] // hs, math.log2(math.e) / math.sqrt(hs)
        for r0, rows in items[: int(n_items[0])].tolist():
            jets = sorted(set(row_jet[r0: r0 + rows].tolist()))
            assert sum(int(jet_off[b + 1] - jet_off[b]) for b in jets) == rows <= 128 and int(jet_off[jets[0]]) == r0
            for b in jets:
                r = slice(int(jet_off[b]), int(jet_off[b + 1]))
                n = r.stop - r.start
                q, k, vv = self._heads(qn, r, n, H, hs), self._heads(kn, r, n, H, hs), self._heads(v, r, n, H, hs)
                sc = q @ k.transpose(1, 2) * c
                m = sc.max(-1, keepdim=True).values
                e = torch.exp2(sc - m)
                inv = 1.0 / e.sum(-1, keepdim=True)
                o[r] = ((e * inv) @ vv).transpose(0, 1).reshape(n, H * hs).to(o.dtype)
                stats[r, :, 0] = m[..., 0].T
                stats[r, :, 1] = inv[..., 0].T

    def attn_tc_bwd(self, dO, qn, kn, v, hs, items, n_items, grid_items, row_jet, jet_off, stats, dqkv):
        C = qn.shape[1]
        H, c = C // hs, math.log2(math.e) / math.sqrt(hs)
        for r0, rows in items[: int(n_items[0])].tolist():
            for b in sorted(set(row_jet[r0: r0 + rows].tolist())):
                r = slice(int(jet_off[b]), int(jet_off[b + 1]))
                n = r.stop - r.start
                q, k, vv, do = (self._heads(t, r, n, H, hs) for t in (qn, kn, v, dO))
                p = torch.exp2(q @ k.transpose(1, 2) * c - stats[r, :, 0].T[..., None]) * stats[r, :, 1].T[..., None]
                dp = do @ vv.transpose(1, 2)
                ds = p * (dp - (p * dp).sum(-1, keepdim=True)) / math.sqrt(hs)
                pb, dsb = p.to(torch.bfloat16).float(), ds.to(torch.bfloat16).float()
                for w, t in enumerate((dsb @ k, dsb.transpose(1, 2) @ q, pb.transpose(1, 2) @ do)):
                    dqkv[r, w * C:(w + 1) * C] = t.transpose(0, 1).reshape(n, C).to(dqkv.dtype)

    def attn_fwd(self, qn, kn, v, jet_off, p_off, B, H, hs, nmax, o, P_, min_n=0):
        for b in range(B):
            r0, r1 = int(jet_off[b]), int(jet_off[b + 1])
            n = r1 - r0
            if n <= min_n:
                continue
            r = slice(r0, r1)
            q, k, vv = self._heads(qn, r, n, H, hs), self._heads(kn, r, n, H, hs), self._heads(v, r, n, H, hs)
            p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hs), -1)
            o[r] = (p @ vv).transpose(0, 1).reshape(n, H * hs).to(o.dtype)
            base = int(p_off[b]) * H
            P_[base: base + H * n * n] = p.to(P_.dtype).flatten()

    def attn_bwd(self, dO, o, P_, qn, kn, v, jet_off, p_off, B, H, hs, nmax, dqkv, C, min_n=0):
        for b in range(B):
            r0, r1 = int(jet_off[b]), int(jet_off[b + 1])
            n = r1 - r0
            if n <= min_n:
                continue
            r = slice(r0, r1)
            q, k, vv, do, oo = (self._heads(t, r, n, H, hs) for t in (qn, kn, v, dO, o))
            base = int(p_off[b]) * H
            p = _f(P_[base: base + H * n * n]).view(H, n, n)
            dv = p.transpose(1, 2) @ do
            delta = (do * oo).sum(-1, keepdim=True)
            ds = p * (do @ vv.transpose(1, 2) - delta) / math.sqrt(hs)
            dq, dk = ds @ k, ds.transpose(1, 2) @ q
            for w, t in enumerate((dq, dk, dv)):
                dqkv[r, w * C:(w + 1) * C] = t.transpose(0, 1).reshape(n, C).to(dqkv.dtype)

    def gelu_fwd(self, z, h):
        h.copy_(F.gelu(_f(z)).to(h.dtype))

    def gelu_bwd(self, dh, z, dz):
        zz = _f(z).clone().requires_grad_(True)
        dz.copy_(torch.autograd.grad(F.gelu(zz), zz, _f(dh))[0].to(dz.dtype))

    def add(self, out, a, y=None, tadd=None, row_jet=None):
        r = a.clone()
        if y is not None:
            r = r + y
        if tadd is not None:
            r = r + tadd[row_jet.long()]
        out.copy_(r)

    def jet_sum(self, g, jet_off, B, out, accumulate=False):
        s = torch.stack([g[int(jet_off[b]):int(jet_off[b + 1])].sum(0) for b in range(B)])
        if accumulate:
            out += s
        else:
            out.copy_(s)

    def head_fwd(self, h, I, wx, bx, wy, by, vt, logits):
        vt.copy_(_f(h[:, :I]) @ wx.T + bx)
        logits.copy_(_f(h[:, I:]) @ wy.T + by)

    def head_bwd(self, dvt, dlog, h, z, I, wx, wy, dz, dwx, dbx, dwy, dby):
        zz = _f(z).clone().requires_grad_(True)
        dh = torch.cat([dvt @ wx, dlog @ wy], 1)
        dz.copy_(torch.autograd.grad(F.gelu(zz), zz, dh)[0].to(dz.dtype))
        dwx += dvt.T @ _f(h[:, :I]); dbx += dvt.sum(0)
        dwy += dlog.T @ _f(h[:, I:]); dby += dlog.sum(0)

    def loss_fwd(self, vt, logits, tgt, k1, jet_off, B, V, loss_mse, loss_ce):
        n = (jet_off[1:] - jet_off[:-1]).float().clamp_min(1)
        rj = torch.repeat_interleave(torch.arange(B), (jet_off[1:] - jet_off[:-1]).long())
        loss_mse.copy_(torch.zeros(B).index_add(0, rj, ((vt - tgt) ** 2).sum(1)) / n)
        loss_ce.copy_(torch.zeros(B).index_add(0, rj, F.cross_entropy(logits, k1.long(), ignore_index=0, reduction="none")) / n)

    def loss_combine(self, loss_mse, loss_ce, u, out5, gl1, gl2, du):
        B = loss_mse.shape[0]
        if u is None:
            out5.copy_(torch.stack([(loss_mse + loss_ce).mean(), loss_mse.mean(), loss_ce.mean(), torch.tensor(1.0), torch.tensor(1.0)]))
            gl1.fill_(1.0 / B); gl2.fill_(1.0 / B)
            return
        w1, w2 = torch.exp(-u[:, 0]), torch.exp(-u[:, 1])
        loss = 0.5 * (u[:, 0] + w1 * loss_mse) + 0.5 * (u[:, 1] + w2 * loss_ce)
        out5.copy_(torch.stack([loss.mean(), loss_mse.mean(), loss_ce.mean(), w1.mean(), w2.mean()]))
        gl1.copy_(0.5 * w1 / B); gl2.copy_(0.5 * w2 / B)
        du[:, 0] = 0.5 * (1 - w1 * loss_mse) / B
        du[:, 1] = 0.5 * (1 - w2 * loss_ce) / B

    def loss_bwd(self, vt, logits, tgt, k1, row_jet, jet_off, gl1, gl2, V, dvt, dlog):
        n = (jet_off[1:] - jet_off[:-1]).float().clamp_min(1)
        rj = row_jet.long()
        real = (torch.arange(vt.shape[0]) < int(jet_off[-1]))[:, None]
        dvt.copy_(torch.where(real, (2 * gl1[rj] / n[rj])[:, None] * (vt - tgt), torch.zeros_like(vt)))
        p = torch.softmax(logits, -1) - F.one_hot(k1.long(), V).float()
        dlog.copy_(torch.where((k1 != 0)[:, None] & real, (gl2[rj] / n[rj])[:, None] * p, torch.zeros_like(p)))

    def sumsq(self, g, out):
        out[0] = (g.double() ** 2).sum().float()

    def adam(self, p, g, m, v, lr, beta1, beta2, eps, step, sumsq=None, max_norm=0.0, grad_scale=1.0, p16=None):
        coef = grad_scale
        if sumsq is not None and max_norm > 0:
            coef *= min(1.0, max_norm / (float(sumsq[0].sqrt()) * grad_scale + 1e-6))
        gi = g * coef
        m.mul_(beta1).add_(gi, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(gi, gi, value=1 - beta2)
        denom = v.sqrt() / math.sqrt(1 - beta2 ** step) + eps
        p.addcdiv_(m, denom, value=-lr / (1 - beta1 ** step))
        if p16 is not None:
            p16.copy_(p.to(torch.bfloat16))
