"""numpy model of the in-place mixed-radix FFT the CUDA kernels use
(csrc/msa_fft.cuh): decimation-in-frequency forward (natural in, digit-reversed
out) and its exact stage-by-stage inverse (digit-reversed in, natural out).
Used by tests to validate the index maps that are baked into the kernels."""
import numpy as np


def dif_forward(x, radices):
    x = np.array(x, dtype=np.complex128)
    N = x.size
    Ns = N
    for R in radices:
        m = Ns // R
        WR = np.exp(-2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)
        for b in range(N // Ns):
            for j in range(m):
                pos = b * Ns + j + m * np.arange(R)
                y = WR @ x[pos]
                x[pos] = y * np.exp(-2j * np.pi * j * np.arange(R) / Ns)
        Ns = m
    return x


def dit_inverse(x, radices):
    """Unnormalised inverse of dif_forward (result is N * original)."""
    x = np.array(x, dtype=np.complex128)
    N = x.size
    sizes = []
    Ns = N
    for R in radices:
        sizes.append((R, Ns))
        Ns //= R
    for R, Ns in reversed(sizes):
        m = Ns // R
        WR = np.exp(+2j * np.pi * np.outer(np.arange(R), np.arange(R)) / R)
        for b in range(N // Ns):
            for j in range(m):
                pos = b * Ns + j + m * np.arange(R)
                v = x[pos] * np.exp(+2j * np.pi * j * np.arange(R) / Ns)
                x[pos] = WR @ v
    return x


def position_of_bin(N, radices):
    """perm[k] = position in the DIF output that holds X[k]."""
    perm = np.zeros(N, dtype=np.int64)
    for p in range(N):
        rem, Ns, k, mult = p, N, 0, 1
        for R in radices:
            m = Ns // R
            d = rem // m
            rem -= d * m
            k += d * mult
            mult *= R
            Ns = m
        perm[k] = p
    return perm
