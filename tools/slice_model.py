"""NumPy model of the arithmetic of the sliced INT8 contraction (csrc/slice_gemm.cu), bit for bit where it matters:

* the power-of-two scale (`pow2_exp`) and its split into two representable factors,
* the balanced byte digits, obtained the way the kernel obtains them (one FP64 addition of 1.5 * 2^52 + 0x80..80 to the
  scaled value, the low mantissa bytes XOR 0x80) and, independently, from the integer definition,
* exact integer slice products with the INT32 range checked, pairs below `smin` dropped,
* FP64 recombination by Horner in powers of 256 and the exact power-of-two unscaling.

It is test infrastructure (tests/test_slice_model.py pins the claims DESIGN.md 4.11 makes about this arithmetic on the CPU);
the product path is the CUDA kernel.
"""
import numpy as np

MAGIC = 6755399441055744.0                       # 1.5 * 2^52


def smin_of(S):
    return 4 if S == 6 else 2


def pow2_exp(amax, S):
    """Largest e with rint(amax * 2^e) representable by S balanced digits (<= 127 * 256^(S-1)); 0 for amax = 0."""
    amax = float(amax)
    if not amax > 0.0:
        return 0
    ex = int(np.frexp(amax)[1])                   # amax = f * 2^ex, f in [0.5, 1)
    e = 8 * S - 1 - ex
    if np.rint(np.ldexp(amax, e)) > np.ldexp(127.0, 8 * S - 8):
        e -= 1
    return e


def split_pow2(e):
    e1 = max(-1000, min(1000, e))
    return np.ldexp(1.0, e1), np.ldexp(1.0, e - e1)


def digits_kernel_way(x, e, S):
    """int8 digits [S, ...] of rint(x * 2^e), extracted from the mantissa of (x * s1) * s2 + magic like the kernel does."""
    x = np.asarray(x, dtype=np.float64)
    s1, s2 = split_pow2(e)
    bias = float(sum(0x80 << (8 * s) for s in range(S - 1)))
    t = (x * s1) * s2 + (MAGIC + bias)            # both products are exact (powers of two): one rounding, as with the DFMA
    bits = t.view(np.uint64)
    out = np.empty((S,) + x.shape, dtype=np.int8)
    for s in range(S):
        byte = ((bits >> np.uint64(8 * s)) & np.uint64(0xFF)).astype(np.uint8)
        if s < S - 1:
            byte = byte ^ np.uint8(0x80)
        out[s] = byte.view(np.int8)
    return out


def digits_definition(v, S):
    """Balanced digits of the Python integers v (object array): v = sum_s d_s 256^s with every d_s in [-128, 127]."""
    v = np.asarray(v, dtype=object)
    out = np.empty((S,) + v.shape, dtype=np.int8)
    flat = v.reshape(-1)
    of = out.reshape(S, -1)
    for idx, val in enumerate(flat):
        val = int(val)
        for s in range(S):
            d = ((val + 128) % 256) - 128 if s < S - 1 else val
            assert -128 <= d <= 127, 'value out of the balanced-digit range'
            of[s, idx] = d
            val = (val - d) // 256
        assert val == 0 or S == 0
    return out


def sliced_matmul(A, B, S=6, smin=None, check_int32=True):
    """C = A @ B (A: M x K, B: K x N) the way slice_gemm_kernel computes it.  Returns (C, info)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    smin = smin_of(S) if smin is None else smin
    ea = pow2_exp(np.abs(A).max() if A.size else 0.0, S)
    eb = [pow2_exp(np.abs(B[:, n]).max() if B.shape[0] else 0.0, S) for n in range(B.shape[1])]
    dA = digits_kernel_way(A, ea, S).astype(np.int64)                              # [S, M, K]
    dB = np.stack([digits_kernel_way(B[:, n], eb[n], S) for n in range(B.shape[1])], axis=2).astype(np.int64)   # [S, K, N]
    nacc = 2 * (S - 1) - smin + 1
    acc = np.zeros((nacc, A.shape[0], B.shape[1]), dtype=np.int64)
    for i in range(S):
        for j in range(S):
            if i + j >= smin:
                acc[i + j - smin] += dA[i] @ dB[j]
    peak = int(np.abs(acc).max()) if acc.size else 0
    if check_int32:
        assert peak < 2 ** 31, 'an INT32 accumulator would overflow'
    r = np.zeros(acc.shape[1:], dtype=np.float64)
    for a in range(nacc - 1, -1, -1):
        r = r * 256.0 + acc[a].astype(np.float64)                                  # the kernel's FMA: same value unless |r| > 2^53
    C = np.empty_like(r)
    for n in range(B.shape[1]):
        E = 8 * smin - ea - eb[n]
        i1, i2 = split_pow2(E)
        C[:, n] = r[:, n] * i1 * i2
    return C, {'ea': ea, 'eb': eb, 'acc_peak': peak, 'nacc': nacc}


def exact_product(A, B):
    """A @ B in exact rational arithmetic (object array of Fractions)."""
    from fractions import Fraction
    Af = np.vectorize(Fraction, otypes=[object])(np.asarray(A, dtype=np.float64))
    Bf = np.vectorize(Fraction, otypes=[object])(np.asarray(B, dtype=np.float64))
    return Af.dot(Bf)
