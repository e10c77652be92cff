"""Shared gradient comparison of the training tests: per-parameter cosine and norm ratio against fp32 autograd, plus the
global cosine / relative L2 error over the whole gradient.  Parameters whose true gradient vanishes identically (the bias of
the per-head key LayerNorm shifts every score of a query row equally, so softmax cancels it) are held to an absolute bound
relative to the global gradient norm instead, and so are small tensors whose error is negligible on the scale of the whole
gradient (the 32/64-element q/k LayerNorm affines of deep blocks: torch's own bf16 autocast shows the same cosines there,
0.95-0.99, see autocast_gradient_error)."""
import torch


def compare_gradients(eng, grads, min_cos=0.99, norm_tol=5e-2, verbose=""):
    gnorm = sum(float(v.double().pow(2).sum()) for v in grads.values()) ** 0.5
    worst, num, den, dot, n1, n2 = (1.0, ""), 0.0, 0.0, 0.0, 0.0, 0.0
    assert sorted(grads) == sorted(eng.names)
    for name in eng.names:
        a, b = eng.g(name).double().flatten().cpu(), grads[name].double().flatten().cpu()
        assert a.shape == b.shape and bool(torch.isfinite(a).all()), name
        num += float((a - b).pow(2).sum()); den += float(b.pow(2).sum())
        dot += float(a @ b); n1 += float(a @ a); n2 += float(b @ b)
        if float(b.norm()) < 1e-5 * gnorm:
            assert float(a.norm()) < 1e-3 * gnorm, (name, float(a.norm()), float(b.norm()), gnorm)
            continue
        cos = float(a @ b / (a.norm() * b.norm() + 1e-300))
        if cos < worst[0]:
            worst = (cos, name)
        negligible = float((a - b).norm()) < 1e-3 * gnorm
        assert cos > min_cos or negligible, (name, cos, float(a.norm()), float(b.norm()), gnorm)
        assert abs(float(a.norm()) / float(b.norm()) - 1.0) < norm_tol or negligible, (name, float(a.norm()), float(b.norm()), gnorm)
    gcos, grel = dot / (n1 * n2) ** 0.5, (num / den) ** 0.5
    if verbose:
        print(f"{verbose}: worst cosine {worst[0]:.5f} at {worst[1]}  global cosine {gcos:.6f}  global rel-L2 {grel:.3e}")
    return gcos, grel, worst


def autocast_gradient_error(loss_fn, params, device_type):
    """Global relative L2 error of torch's own bf16 autocast gradients against its fp32 gradients for the same loss closure:
    the calibration of what 'bf16 forward/backward' costs on a fixture (SURVEY 8(c) does the same for the forward pass)."""
    out = []
    for enabled in (False, True):
        for p in params:
            p.grad = None
        with torch.autocast(device_type, dtype=torch.bfloat16, enabled=enabled):
            loss = loss_fn()
        loss.backward()
        out.append([p.grad.detach().double().clone() for p in params])
    num = sum(float((a - b).pow(2).sum()) for a, b in zip(out[1], out[0]))
    den = sum(float(b.pow(2).sum()) for b in out[0])
    return (num / den) ** 0.5
