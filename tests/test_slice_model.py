"""The arithmetic of the opt-in sliced INT8 contraction, pinned on the CPU with the NumPy model
tools/slice_model.py: digits, ranges, truncation and error bound -- the claims the CUDA kernel (csrc/slice_gemm.cu, tested
on the GPU in tests/test_gpu_slices.py) relies on."""
import os
import sys
from fractions import Fraction

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'tools'))
import slice_model as sm   # noqa: E402


@pytest.mark.parametrize('S', [5, 6])
def test_kernel_digit_extraction_equals_the_integer_definition(S):
    rng = np.random.RandomState(S)
    x = np.concatenate([rng.randn(2000), rng.randn(200) * 1e-6, [0.0, -0.0, 1.0, -1.0, 0.9921875, -0.9921875, 5.8, -5.8]])
    x = x / np.abs(x).max() * 3.7
    e = sm.pow2_exp(np.abs(x).max(), S)
    d = sm.digits_kernel_way(x, e, S)
    v = [int(np.rint(np.ldexp(float(t), e))) for t in x]
    assert max(abs(t) for t in v) <= 127 * 256 ** (S - 1)
    assert np.array_equal(d, sm.digits_definition(np.array(v, dtype=object), S))
    # the digits ARE the number: sum d_s 256^s = rint(x 2^e), every digit a signed byte
    back = sum(d[s].astype(object) * (256 ** s) for s in range(S))
    assert list(back) == v


@pytest.mark.parametrize('S', [5, 6])
@pytest.mark.parametrize('amax', [1.0, 0.9921875, 0.9921875 * (1 + 2.0 ** -40), 1 - 2.0 ** -52, 2.0 ** -1060, 1.7e308, 3.0, 1e-300])
def test_scale_keeps_the_largest_value_inside_the_digit_range(S, amax):
    e = sm.pow2_exp(amax, S)
    v = int(np.rint(np.ldexp(amax, e)))
    assert v <= 127 * 256 ** (S - 1)
    if amax >= 2.0 ** -1000:
        assert v >= 127 * 256 ** (S - 1) // 4                     # and wastes less than two bits
    s1, s2 = sm.split_pow2(e)
    assert s1 > 0 and s2 > 0 and np.isfinite(s1) and np.isfinite(s2)
    assert np.array_equal(sm.digits_kernel_way(np.array([amax, -amax]), e, S), sm.digits_definition(np.array([v, -v], dtype=object), S))


def test_int32_accumulators_hold_a_whole_segment():
    # worst case of one K segment: 128 chunks x 128 bytes = 16384 products, every digit at -128, six pairs in one accumulator
    assert 6 * 16384 * 128 * 128 < 2 ** 31
    S, K = 6, 16384
    A = np.full((1, K), -1.0)                                     # -2^46 -> digits (-64, 0, 0, 0, 0, 0) ... use the model for the real peak
    B = np.full((K, 1), -1.0)
    C, info = sm.sliced_matmul(A, B, S)
    assert C[0, 0] == float(K) and info['acc_peak'] < 2 ** 31


@pytest.mark.parametrize('S', [5, 6])
def test_dropped_pairs_are_below_2_to_the_minus_52_of_full_scale(S):
    smin = sm.smin_of(S)
    K = 16384
    dropped = sum(128 * 128 * K * 256 ** (i + j) for i in range(S) for j in range(S) if i + j < smin)
    full = K * (127 * 256 ** (S - 1)) ** 2
    assert Fraction(dropped, full) < Fraction(1, 2 ** 52) * 40       # 2^-52 .. 2^-47 of the largest possible product sum
    # relative to the rounding of the operands themselves (2^-(8S-1) each) it is negligible
    assert Fraction(dropped, full) < Fraction(1, 2 ** (8 * S - 1)) / 8


@pytest.mark.parametrize('S,dist', [(6, 'uniform'), (6, 'normal'), (5, 'uniform'), (5, 'normal'), (6, 'wide')])
def test_model_error_stays_inside_the_stated_bound(S, dist):
    rng = np.random.RandomState(7)
    M, K, N = 12, 96, 5
    if dist == 'uniform':
        A, B = rng.rand(M, K), rng.rand(K, N)
    elif dist == 'normal':
        A, B = rng.randn(M, K), rng.randn(K, N)
    else:
        A = rng.randn(M, K) * 10.0 ** rng.uniform(-3, 3, size=(M, K))
        B = rng.randn(K, N) * 10.0 ** rng.uniform(-3, 3, size=(K, N))
    C, info = sm.sliced_matmul(A, B, S)
    exact = sm.exact_product(A, B)
    err = np.array([[abs(Fraction(float(C[m, n])) - exact[m, n]) for n in range(N)] for m in range(M)], dtype=object)
    bound = Fraction(1, 2 ** (8 * S - 1)) * (Fraction(float(np.abs(A).max())) * Fraction(float(np.abs(B).sum(axis=0).max())) +
                                              Fraction(float(np.abs(B).max())) * Fraction(float(np.abs(A).sum(axis=1).max())))
    assert max(err.reshape(-1)) <= 2 * bound + Fraction(float(np.abs(C).max())) / 2 ** 51
    if S == 6 and dist != 'wide':
        assert float(max(err.reshape(-1))) <= 1e-13 * float(np.abs(C).max())


def test_exact_on_small_integers_and_extreme_magnitudes():
    rng = np.random.RandomState(1)
    A = rng.randint(-1000, 1000, size=(7, 40)).astype(float)
    B = rng.randint(-50, 50, size=(40, 3)).astype(float)
    C, _ = sm.sliced_matmul(A, B, 6)
    assert np.array_equal(C, A.dot(B))
    for mag in (1e-300, 1e300 / 64):
        C, _ = sm.sliced_matmul(A * mag / 1000.0, B, 6)
        ref = (A * mag / 1000.0).dot(B)
        assert np.abs(C - ref).max() <= 1e-12 * np.abs(ref).max()
