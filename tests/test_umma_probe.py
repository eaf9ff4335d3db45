"""Pins every tcgen05.mma operand form (shared-memory / TMEM layouts and descriptor fields) the fused
kernels use, one small MMA chain at a time, against a plain matmul."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

KSW, TMEM, MNSW, KNONE, MNNONE = 0, 1, 2, 3, 4


def _probe(a_mode, b_mode, N, ksteps, a_lbo, a_sbo, b_lbo, b_sbo, a_off, b_off, seed=0):
    from egt_b200 import _lib as L
    lib = L.load()
    K = 16 * ksteps
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(128, K, generator=g).bfloat16().to(DEV)
    B = torch.randn(N, K, generator=g).bfloat16().to(DEV)
    D = torch.full((128, N), float('nan'), device=DEV)
    ao = (C.c_uint32 * 8)(*(list(a_off) + [0] * (8 - len(a_off))))
    bo = (C.c_uint32 * 8)(*(list(b_off) + [0] * (8 - len(b_off))))
    L.check(lib.egt_debug_umma_probe(a_mode, b_mode, N, ksteps, a_lbo, a_sbo, b_lbo, b_sbo, ao, bo,
                                     C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    torch.testing.assert_close(D, ref, rtol=1e-4, atol=1e-3)


def test_f1_ss_kmajor_sw128_both():
    """S = Q * Kexp^T: A and B K-major, 128B swizzle, k-step = +32 bytes."""
    _probe(KSW, KSW, 16, 4, 16, 1024, 16, 1024, [0, 32, 64, 96], [0, 32, 64, 96])


def test_f1b_ss_kmajor_sw128_k128():
    _probe(KSW, KSW, 16, 8, 16, 1024, 16, 1024, [0, 32, 64, 96, 16384, 16416, 16448, 16480],
           [0, 32, 64, 96, 16384, 16416, 16448, 16480])


@pytest.mark.parametrize('N,ksteps', [(32, 1), (32, 4), (16, 1)])
def test_f2_ss_a_sw128_b_kmajor_none(N, ksteps):
    """E|G = e_tile * Wblk^T: B is a small K-major matrix of un-swizzled 8x16B core matrices."""
    lbo = N * 16
    _probe(KSW, KNONE, N, ksteps, 16, 1024, lbo, 128, [32 * s for s in range(ksteps)],
           [2 * s * lbo for s in range(ksteps)])


@pytest.mark.parametrize('N,ksteps', [(64, 1), (64, 2), (128, 1)])
def test_f3_ts_b_mnmajor_sw128(N, ksteps):
    """O += A~ * Vexp: A from TMEM (packed bf16), B MN-major with 128B swizzle (rows = k, 128 bytes of n)."""
    K = 16 * ksteps
    _probe(TMEM, MNSW, N, ksteps, 0, 0, K * 128, 1024, [8 * s for s in range(ksteps)],
           [2 * s * 1024 for s in range(ksteps)])


@pytest.mark.parametrize('N,ksteps', [(16, 1), (16, 2)])
def test_f4_ts_b_kmajor_none(N, ksteps):
    """de' = H^ * Wr_blk: A from TMEM, B small K-major un-swizzled."""
    lbo = N * 16
    _probe(TMEM, KNONE, N, ksteps, 0, 0, lbo, 128, [8 * s for s in range(ksteps)], [2 * s * lbo for s in range(ksteps)])


@pytest.mark.parametrize('N', [16, 32])
def test_f5_ss_a_mnmajor_sw128_b_mnmajor_none(N):
    """dK|dV (transposed) = [dO;Q]^T-form * B: A MN-major 128B swizzle (M = 2 atoms of 64), B MN-major
    un-swizzled written row by row (k = query row), K = 128."""
    ksteps, K = 8, 128
    _probe(MNSW, MNNONE, N, ksteps, K * 128, 1024, 128, (K // 8) * 128, [2 * s * 1024 for s in range(ksteps)],
           [2 * s * 128 for s in range(ksteps)])


def test_f6_ts_b_kmajor_sw128():
    """dQ-style chain with B K-major swizzled and A from TMEM."""
    _probe(TMEM, KSW, 64, 2, 0, 0, 16, 1024, [0, 8], [0, 32])


# ---- operand forms added by the width-generic kernels (wide_fwd.cu / wide_bwd.cu) -------------------------
@pytest.mark.parametrize('N,ksteps', [(32, 2), (16, 4), (16, 2), (64, 2)])
def test_w1_ss_a_sw128_b_kmajor_none_wide(N, ksteps):
    """[E|G] of one key at d_e = 32 / 64 (K = d_e), dH_ext = de' W_r^T (N = 16)."""
    lbo = N * 16
    _probe(KSW, KNONE, N, ksteps, 16, 1024, lbo, 128, [32 * s for s in range(ksteps)],
           [2 * s * lbo for s in range(ksteps)])


@pytest.mark.parametrize('N', [32, 64])
def test_w2_ts_b_kmajor_none_wide(N):
    """e' += H^ W_r with N = d_e."""
    lbo = N * 16
    _probe(TMEM, KNONE, N, 1, 0, 0, lbo, 128, [0], [0])


@pytest.mark.parametrize('N', [48, 96, 128])
def test_w3_ts_b_mnmajor_sw128_wide(N):
    """O += A~ Vexp with N = d in {48, 96, 128}: B is 16 rows of 128-byte swizzled atoms, LBO = 2048."""
    _probe(TMEM, MNSW, N, 1, 0, 0, 2048, 1024, [0], [0])


def test_w4_ss_a_kmajor_none():
    """bias product: A is an un-swizzled K-major [128 x 16] tile in shared memory (LBO = 2048, SBO = 128)."""
    _probe(KNONE, KNONE, 32, 1, 2048, 128, 32 * 16, 128, [0], [0])


def test_w5_ss_kmajor_sw128_k96():
    """S = Q Kexp^T at d = 96: six k-steps over one and a half swizzle atoms; B atoms are 2048 bytes apart."""
    a_off = [0, 32, 64, 96, 16384, 16416]
    b_off = [0, 32, 64, 96, 2048, 2080]
    _probe_b16(a_off, b_off)


def _probe_b16(a_off, b_off):
    """K-major swizzled B with only 16 rows per atom (atom stride 2048 instead of the probe's 16384 image)."""
    from egt_b200 import _lib as L
    lib = L.load()
    ksteps = len(a_off)
    K = 16 * ksteps
    g = torch.Generator().manual_seed(1)
    A = torch.randn(128, K, generator=g).bfloat16().to(DEV)
    B = torch.randn(16, K, generator=g).bfloat16().to(DEV)
    D = torch.full((128, 16), float('nan'), device=DEV)
    ao = (C.c_uint32 * 8)(*(list(a_off) + [0] * (8 - ksteps)))
    bo = (C.c_uint32 * 8)(*(list(b_off) + [0] * (8 - ksteps)))
    # b_mode 5: K-major 128B swizzle with 2048-byte atoms (16 rows)
    L.check(lib.egt_debug_umma_probe(KSW, 5, 16, ksteps, 16, 1024, 16, 1024, ao, bo, C.c_void_p(A.data_ptr()),
                                     C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    torch.testing.assert_close(D, A.float() @ B.float().t(), rtol=1e-4, atol=1e-3)


def test_w6_ss_a_mnmajor_none_b_mnmajor_sw128():
    """weight-gradient products at h = 8: A is an un-swizzled MN-major image written row by row (k = query row),
    B the raw edge tile as it lies in the TMA stage (MN-major, 128B swizzle), K = 128."""
    ksteps, K = 8, 128
    _probe(MNNONE, MNSW, 32, ksteps, 128, (K // 8) * 128, K * 128, 1024, [2 * s * 128 for s in range(ksteps)],
           [2 * s * 1024 for s in range(ksteps)])
