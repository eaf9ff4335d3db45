"""numpy replica of the kernels' counter RNG (egt_b200/csrc/common.cuh: philox4x32_10 / rng_uniform).

Test infrastructure: lets the CPU oracle see exactly the uniform draws the CUDA kernels use for the
random key mask (stream 0) and attention dropout (stream 1), so mask parity can be bit-exact."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def uniform(seed, offset, stream_id, idx):
    """u in (0,1) for flat element indices idx = ((b*N+l)*N+m)*h+hh  (array of uint64)."""
    idx = np.asarray(idx, dtype=np.uint64)
    q = idx >> np.uint64(3)
    lane = (idx & np.uint64(7)).astype(np.int64)
    c0 = q & MASK32
    c1 = q >> np.uint64(32)
    c2 = np.full_like(q, offset & 0xFFFFFFFF)
    c3 = np.full_like(q, ((offset >> 32) + stream_id * 0x40000000) & 0xFFFFFFFF)
    r = philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(r, axis=-1)                                   # [..., 4]
    word = np.take_along_axis(words, (lane >> 1)[..., None], axis=-1)[..., 0]
    bits = np.where(lane & 1, word >> np.uint64(16), word & np.uint64(0xFFFF))
    return ((bits.astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 65536.0)).astype(np.float32)


def elem_index(B, N, h):
    """RNG element index of every (b, l, m, hh) (common.cuh: rng_elem_index): one Philox call = 2 consecutive
    keys x 4 consecutive heads."""
    b, l, m, hh = np.meshgrid(np.arange(B, dtype=np.uint64), np.arange(N, dtype=np.uint64), np.arange(N, dtype=np.uint64),
                              np.arange(h, dtype=np.uint64), indexing='ij')
    npair, nq = np.uint64((N + 1) // 2), np.uint64((h + 3) // 4)
    u = np.uint64
    return ((((b * u(N) + l) * npair + (m >> u(1))) * nq + (hh >> u(2))) << u(3)) | ((m & u(1)) << u(2)) | (hh & u(3))


def noise_tensor(seed, offset, stream_id, B, N, h):
    return uniform(seed, offset, stream_id, elem_index(B, N, h))
