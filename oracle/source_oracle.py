"""TEST INFRASTRUCTURE - CPU restatement of the sampler's source construction (never imported by the product package).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.

The LAW is the reference's:
  utils/aoj.py:875-890          sample_from_empirical_masks: nums = mask.sum(1); probs = density histogram over
                                bins arange(0, D + 2); multiplicity ~ Categorical(probs); mask[i, :n] = 1
  scripts/sample_mmf.py:82-84   noise_continuous = randn_like(continuous) * pad_mask
                                noise_discrete = randint_like(discrete, 1, vocab_size) * pad_mask
The reference draws from torch's global generators, which no other implementation can reproduce, so the DRAWS are defined
here as counter-based functions of (seed, global jet, slot) - Philox4x32-10 (Salmon et al., SC'11; checked below against the
Random123 known-answer vectors), inverse-CDF for the multiplicity, Box-Muller for the normals:
  jet J:   w = Philox(ctr = (J lo, J hi, 0, 'MULT'), key = seed);  u = (w0 >> 8) 2^-24;  n = #{m < D : cdf[m] <= u}
  slot S = J D + d (d < n):  w = Philox(ctr = (S lo, S hi, 0, 'SRCE'), key = seed);  u_i = ((w_i >> 9) + 0.5) 2^-23
           x = (r0 cos 2 pi u1, r0 sin 2 pi u1, r2 cos 2 pi u3),  r_i = sqrt(-2 ln u_i)
           k = 1 + ((low bytes of w0..w3 as one 32-bit word) * (V - 1) >> 32)
Masks, multiplicities and tokens are integer work: the CUDA kernel must match bit for bit.  Normals are compared to 1e-5.
Parity with the reference itself is in distribution (tests/test_source.py: histogram law against the reference's own
sample_from_empirical_masks run on the same empirical masks, moments / KS of the normals, uniformity of the tokens).
"""
from __future__ import annotations

import numpy as np

TAG_MULT, TAG_SLOT = 0x4D554C54, 0x53524345
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10; counters are uint32 arrays (or scalars), key two python ints. Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _MASK32 for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & _MASK32, p1 >> np.uint64(32), p1 & _MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return tuple(np.asarray(c, dtype=np.uint64).astype(np.uint32) for c in (c0, c1, c2, c3))


def multiplicity_cdf(mult_probs) -> np.ndarray:
    """cdf[m] = P(multiplicity <= m) accumulated in double and rounded once to float32; the last entry is exactly 1."""
    p = np.asarray(mult_probs, dtype=np.float32).astype(np.float64)
    cdf = (np.cumsum(p) / p.sum()).astype(np.float32)
    cdf[-1] = np.float32(1.0)
    return cdf


def empirical_multiplicity_probs(pad_masks: np.ndarray, max_num_particles: int) -> np.ndarray:
    """utils/aoj.py:876-878."""
    nums = np.asarray(pad_masks).reshape(len(pad_masks), -1).sum(1)
    probs, _ = np.histogram(nums, bins=np.arange(0, max_num_particles + 2, 1), density=True)
    return probs.astype(np.float32)


def make_source(mult_probs, num_jets: int, max_num_particles: int, vocab_size: int, seed: int, first_global_jet: int = 0):
    """Returns x0 (B,D,3) f32, k0 (B,D) i64, mask (B,D) i64, n (B,) i32."""
    B, D, V = int(num_jets), int(max_num_particles), int(vocab_size)
    k0key, k1key = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    cdf = multiplicity_cdf(mult_probs)
    J = np.arange(B, dtype=np.uint64) + np.uint64(first_global_jet)
    w = philox4x32_10(J & _MASK32, J >> np.uint64(32), 0, TAG_MULT, k0key, k1key)
    u = (w[0] >> np.uint32(8)).astype(np.float32) * np.float32(5.96046448e-08)
    n = (cdf[None, :D] <= u[:, None]).sum(1).astype(np.int32)
    d = np.arange(D, dtype=np.uint64)
    S = J[:, None] * np.uint64(D) + d[None, :]
    w = philox4x32_10(S & _MASK32, S >> np.uint64(32), 0, TAG_SLOT, k0key, k1key)
    uu = [((wi >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.1920929e-07) for wi in w]
    u64 = [x.astype(np.float64) for x in uu]
    r0, r2 = np.sqrt(-2.0 * np.log(u64[0])), np.sqrt(-2.0 * np.log(u64[2]))
    x = np.stack([r0 * np.cos(2 * np.pi * u64[1]), r0 * np.sin(2 * np.pi * u64[1]), r2 * np.cos(2 * np.pi * u64[3])], axis=-1)
    bits = ((w[0] & np.uint32(0xFF)) | ((w[1] & np.uint32(0xFF)) << np.uint32(8)) | ((w[2] & np.uint32(0xFF)) << np.uint32(16))
            | ((w[3] & np.uint32(0xFF)) << np.uint32(24))).astype(np.uint64)
    k = (1 + ((bits * np.uint64(V - 1)) >> np.uint64(32))).astype(np.int64)
    mask = (np.arange(D)[None, :] < n[:, None]).astype(np.int64)
    return (x * mask[..., None]).astype(np.float32), k * mask, mask, n
