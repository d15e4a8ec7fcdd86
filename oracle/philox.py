""" TEST INFRASTRUCTURE ONLY — numpy restatement of the device collocation sampler.

The reference samples on the host (`torch.rand((B,1))` per column, pydens/model_torch.py:431, or a
batchflow `NumpySampler`, :433); the engine replaces that with a counter-based generator evaluated in
the kernel (pydens_b200/csrc/pinn_device.cuh: philox4x32_10, sample_column).  RNG streams of the
reference are unspecified, so parity with it is distribution-level only ("parity unpinned" for the
stream); what IS pinned, bit-exactly, is this restatement against (a) the published Philox4x32-10
known-answer vectors (Salmon et al., SC'11, Random123 kat_vectors) and (b) the device output.

Counter = (gidx lo, gidx hi, step lo, (step hi & 0xffff) << 16 | block), key = (seed lo, seed hi).
Uniform column k < 4 uses word k of block 0 (k >= 4: word k-4 of block 1): lo + (hi-lo) * (w >> 8) * 2^-24.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """ Vectorised over numpy uint32 arrays. Returns 4 uint32 arrays. """
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def sample(cols, total, seed, step, point_offset, n):
    """ cols: [(kind, a, b)] per column (kind 0 uniform [a,b), 1 normal(a, b), 2 const a) — or
    ('mix', group key, [(weight, kind, a, b), ...]) for a mixture column — or None for the default U[0,1).  Returns float32 [n, total] — bit-exact with pinn_sample for uniform / const
    columns (normal columns go through libm log/cos and are compared with a tolerance). """
    if cols is None:
        cols = [(0, 0.0, 1.0)] * total
    gidx = np.uint64(point_offset) + np.arange(n, dtype=np.uint64)
    c0 = (gidx & MASK).astype(np.uint32)
    c1 = (gidx >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(n, step & 0xFFFFFFFF, dtype=np.uint32)
    c3base = np.uint32(((step >> 32) & 0xFFFF) << 16)
    k0, k1 = np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)
    blocks = {}

    def block(b):
        if b not in blocks:
            blocks[b] = philox4x32_10(c0, c1, c2, np.full(n, c3base | np.uint32(b), dtype=np.uint32), k0, k1)
        return blocks[b]

    out = np.zeros((n, total), dtype=np.float32)
    scale = np.float32(2.0 ** -24)

    def simple(k, kind, a, b):
        a32, b32 = np.float32(a), np.float32(b)
        if kind == 2:
            return np.full(n, a32, dtype=np.float32)
        if kind == 0:
            w = block(0)[k] if k < 4 else block(1)[k - 4]
            u = (w >> np.uint32(8)).astype(np.float32) * scale
            # fmaf(b - a, u, a): one rounding — emulate in float64 (exact product of two fp32 fits)
            return ((np.float64(b32 - a32)) * u.astype(np.float64) + np.float64(a32)).astype(np.float32)
        w = block(2 + k)
        u1 = ((w[0] >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * scale
        u2 = (w[1] >> np.uint32(8)).astype(np.float32) * scale
        rad = np.sqrt(np.float32(-2.0) * np.log(u1))
        z = rad * np.cos(np.float32(6.283185307179586) * u2)
        return (np.float64(b32) * z.astype(np.float64) + np.float64(a32)).astype(np.float32)

    def tnormal(k, a, b, lo, hi):
        # restatement of sample_tnormal (pinn_device.cuh): candidates from words (0,1) / (2,3) of block 2+k with the
        # attempt number in bits 8..15 of the block word; first hit wins; after 16 misses the clamped mean
        a32, b32, lo32, hi32 = np.float32(a), np.float32(b), np.float32(lo), np.float32(hi)
        res = np.full(n, np.minimum(np.maximum(a32, lo32), hi32), dtype=np.float32)
        todo = np.ones(n, dtype=bool)
        for att in range(8):
            w = philox4x32_10(c0, c1, c2, np.full(n, c3base | np.uint32(att << 8) | np.uint32(2 + k), dtype=np.uint32), k0, k1)
            for half in range(2):
                w1, w2 = (w[2], w[3]) if half else (w[0], w[1])
                u1 = ((w1 >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * scale
                u2 = (w2 >> np.uint32(8)).astype(np.float32) * scale
                z = np.sqrt(np.float32(-2.0) * np.log(u1)) * np.cos(np.float32(6.283185307179586) * u2)
                v = (np.float64(b32) * z.astype(np.float64) + np.float64(a32)).astype(np.float32)
                hit = todo & (v >= lo32) & (v <= hi32)
                res[hit] = v[hit]
                todo &= ~hit
        return res

    groups = {}
    for k, col in enumerate(cols):
        if col[0] == 'mix':                        # ('mix', group key, [(weight, kind, a, b), ...])
            _, key, comps = col
            g = groups.setdefault(key, len(groups))
            u = (block(10 + g)[0] >> np.uint32(8)).astype(np.float32) * scale
            total_w, acc, cum = float(sum(c[0] for c in comps)), 0.0, []
            for j, c in enumerate(comps):
                acc += float(c[0])
                cum.append(np.float32(1.0 if j == len(comps) - 1 else acc / total_w))
            idx = np.zeros(n, dtype=np.int64)
            for c in range(len(comps) - 1):
                idx += (u >= cum[c]).astype(np.int64)
            vals = np.stack([simple(k, kind, a, b) for _, kind, a, b in comps], axis=0)
            out[:, k] = vals[idx, np.arange(n)]
        elif len(col) == 5:                        # (4, loc, scale, low, high): truncated normal
            out[:, k] = tnormal(k, *col[1:])
        else:
            out[:, k] = simple(k, *col)
    return out
