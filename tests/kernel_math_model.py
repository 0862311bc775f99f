"""numpy model of the index/twiddle math the CUDA kernels implement (test infrastructure).

Each function mirrors one device-side stage in ndrustfft_b200/csrc/tile_kernels.cuh:
  * `stockham_fft`     — the autosort radix passes (one ping-pong pass per radix)
  * `bluestein_fft`    — chirp-z wrapper for lengths with large prime factors
  * `pro_*` / `epi_*`  — the prologue/epilogue of every transform kind around a forward complex FFT core
`transform(kind, x, n)` composes them exactly like the kernel does; tests/test_kernel_math_model.py checks
it against the oracle, so a formula error shows up on the CPU before any GPU time is spent.
All results are "Default"-free: unscaled engine outputs (rustfft / realfft / rustdct conventions).
"""
import numpy as np


def factorize(N, radices=(16, 8, 4, 2, 3, 5, 7, 11, 13)):
    """Greedy radix schedule; returns None if N has a prime factor > 13."""
    out = []
    rem = N
    for r in (16, 8, 4, 2):
        while rem % r == 0:
            out.append(r)
            rem //= r
    for r in (3, 5, 7, 11, 13):
        while rem % r == 0:
            out.append(r)
            rem //= r
    return out if rem == 1 else None


def stockham_fft(x, radices):
    """Forward DFT of x (len N = prod(radices)) by Stockham autosort passes.

    Pass with radix r and p = product of previous radices:  for i in [0, N/r): k = i % p,
      u[q] = x[i + q*N/r] * W_N^{q*k*N/(p*r)},  v = DFT_r(u),  y[(i-k)*r + k + q*p] = v[q]."""
    x = np.asarray(x, dtype=np.complex128).copy()
    N = x.shape[0]
    assert int(np.prod(radices)) == N
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    p = 1
    for r in radices:
        t = N // r
        y = np.empty_like(x)
        step = N // (p * r)
        dft = np.exp(-2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
        for i in range(t):
            k = i % p
            u = np.array([x[i + q * t] * tw[(q * k * step) % N] for q in range(r)])
            v = dft @ u
            base = (i - k) * r + k
            for q in range(r):
                y[base + q * p] = v[q]
        x = y
        p *= r
    return x


def next_pow2(v):
    m = 1
    while m < v:
        m *= 2
    return m


def bluestein_fft(z):
    """Forward DFT of arbitrary length N through a power-of-two convolution of length M >= 2N-1."""
    z = np.asarray(z, dtype=np.complex128)
    N = z.shape[0]
    M = next_pow2(2 * N - 1)
    j = np.arange(N)
    c = np.exp(-1j * np.pi * ((j * j) % (2 * N)) / N)           # chirp, exact phase reduction
    b = np.zeros(M, np.complex128)
    b[:N] = np.conj(c)
    b[M - N + 1:] = np.conj(c[1:][::-1])
    bhat = stockham_fft(b, factorize(M))                          # precomputed on the host in the plan
    a = np.zeros(M, np.complex128)
    a[:N] = z * c
    A = stockham_fft(a, factorize(M))
    P = A * bhat
    conv = np.conj(stockham_fft(np.conj(P), factorize(M))) / M    # inverse via conj-forward-conj
    return conv[:N] * c


def core_fft(z):
    N = len(z)
    if N == 1:
        return np.asarray(z, dtype=np.complex128).copy()
    f = factorize(N)
    return stockham_fft(z, f) if f is not None else bluestein_fft(z)


# ---------------- kinds ----------------
def c2c(x, inverse):
    if not inverse:
        return core_fft(x)
    return np.conj(core_fft(np.conj(x)))


def r2c(x):
    n = len(x)
    m = n // 2 + 1
    if n % 2 == 1 or n < 2:
        return core_fft(np.asarray(x, dtype=np.complex128))[:m]
    N = n // 2
    z = x[0::2] + 1j * x[1::2]
    Z = core_fft(z)
    return r2c_post(Z, n)


def r2c_post(Z, n):
    """X[k], k=0..N from Z = FFT_N(x[2j] + i x[2j+1]);  w_k = exp(-2 pi i k / n)."""
    N = n // 2
    k = np.arange(N + 1)
    Zk = Z[k % N]
    Zc = np.conj(Z[(N - k) % N])
    w = np.exp(-2j * np.pi * k / n)
    return 0.5 * (Zk + Zc) - 0.5j * w * (Zk - Zc)


def c2r_pre(X, n):
    """Z[k], k=0..N-1 such that conj(FFT_N(conj Z)) = x[2j] + i x[2j+1] with x = unscaled inverse real DFT."""
    N = n // 2
    k = np.arange(N)
    Xk = X[k]
    Xc = np.conj(X[N - k])
    w = np.exp(+2j * np.pi * k / n)
    return (Xk + Xc) + 1j * w * (Xk - Xc)


def c2r(X, n):
    X = np.array(X, dtype=np.complex128)
    X[0] = X[0].real
    if n % 2 == 0:
        X[n // 2] = X[n // 2].real
    if n % 2 == 1 or n < 2:
        full = np.zeros(n, np.complex128)
        m = n // 2 + 1
        full[:m] = X[:m]
        for k in range(1, m):
            full[n - k] = np.conj(X[k])
        return c2c(full, True).real
    Z = c2r_pre(X, n)
    z = c2c(Z, True)
    out = np.empty(n)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out


def makhoul(x):
    n = len(x)
    v = np.empty(n, dtype=np.asarray(x).dtype)
    h = (n + 1) // 2
    v[:h] = x[0::2]
    v[h:] = x[1::2][::-1]
    return v


def dct2(x):
    """rustdct DCT-II: sum_j x_j cos(pi k (2j+1) / (2n))."""
    n = len(x)
    v = makhoul(np.asarray(x, dtype=np.float64))
    k = np.arange(n)
    t = np.exp(-1j * np.pi * k / (2 * n))
    if n % 2 == 1 or n < 2:
        V = core_fft(v.astype(np.complex128))
        return (V * t).real
    N = n // 2
    V = r2c(v)                       # k = 0..N
    A = V * t[: N + 1]
    y = np.empty(n)
    y[: N + 1] = A.real
    kk = np.arange(1, N)
    y[n - kk] = -A[kk].imag
    return y


def dct3(y):
    """rustdct DCT-III: y_0/2 + sum_{k>=1} y_k cos(pi k (2j+1) / (2n))."""
    n = len(y)
    y = np.asarray(y, dtype=np.float64)
    yy = np.concatenate([y, [0.0]])
    k = np.arange(n)
    tc = np.exp(+1j * np.pi * k / (2 * n))
    if n % 2 == 1 or n < 2:
        V = tc * (yy[k] - 1j * yy[n - k])
        v = c2c(V, True).real * 0.5
    else:
        N = n // 2
        kk = np.arange(N + 1)
        V = tc[: N + 1] * (yy[kk] - 1j * yy[n - kk])
        v = c2r(V, n) * 0.5
    x = np.empty(n)
    h = (n + 1) // 2
    x[0::2] = v[:h]
    x[1::2] = v[h:][::-1]
    return x


def dct1(x):
    """rustdct DCT-I: x_0/2 + (-1)^k x_{n-1}/2 + sum_{j=1}^{n-2} x_j cos(pi j k/(n-1)); n >= 2."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    N = n - 1
    # even extension e of length 2N: e[j] = x[j] (j<=N), e[2N-j] = x[j]; pack z[j] = e[2j] + i e[2j+1]
    e = np.concatenate([x, x[-2:0:-1]])
    z = e[0::2] + 1j * e[1::2]
    Z = core_fft(z)
    Y = r2c_post(Z, 2 * N)
    return 0.5 * Y.real


def dct4(x):
    """rustdct DCT-IV: sum_j x_j cos(pi (2j+1)(2k+1)/(4n))."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    if n % 2 == 1:
        # odd n: zero-padded length-2n complex FFT with pre/post twiddles
        j = np.arange(n)
        a = np.zeros(2 * n, np.complex128)
        a[:n] = x * np.exp(-1j * np.pi * j / (2 * n))
        A = core_fft(a)
        k = np.arange(n)
        return (np.exp(-1j * np.pi * (2 * k + 1) / (4 * n)) * A[:n]).real
    N = n // 2
    j = np.arange(N)
    u = (x[2 * j] + 1j * x[n - 1 - 2 * j]) * np.exp(-1j * np.pi * j / n)
    U = core_fft(u)
    C = U * np.exp(-1j * np.pi * (4 * j + 1) / (4 * n))
    y = np.empty(n)
    y[2 * j] = C.real
    y[n - 1 - 2 * j] = -C.imag
    return y


# ---------------- in-place decimation passes (what tile_kernels.cuh actually runs) ----------------
def dif_inplace(x, radices):
    """In-place decimation-in-frequency passes.  Result is left in mixed-radix digit-reversed
    positions: X[k] sits at position perm(radices)[k]."""
    x = np.asarray(x, dtype=np.complex128).copy()
    N = len(x)
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    ns = N
    for r in radices:
        s = ns // r
        dft = np.exp(-2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
        for b in range(N // r):
            blk, j = divmod(b, s)
            base = blk * ns + j
            u = np.array([x[base + q * s] for q in range(r)])
            v = dft @ u
            for q in range(r):
                x[base + q * s] = v[q] * tw[q * j * (N // ns)]
        ns = s
    return x


def dit_inplace(x, radices):
    """Transpose of `dif_inplace`: takes digit-reversed input, yields natural-order DFT."""
    x = np.asarray(x, dtype=np.complex128).copy()
    N = len(x)
    tw = np.exp(-2j * np.pi * np.arange(N) / N)
    sub = []
    ns = N
    for r in radices:
        sub.append((r, ns))
        ns //= r
    for r, ns in reversed(sub):
        s = ns // r
        dft = np.exp(-2j * np.pi * np.outer(np.arange(r), np.arange(r)) / r)
        for b in range(N // r):
            blk, j = divmod(b, s)
            base = blk * ns + j
            u = np.array([x[base + q * s] * tw[q * j * (N // ns)] for q in range(r)])
            v = dft @ u
            for q in range(r):
                x[base + q * s] = v[q]
    return x


def perm(radices):
    """perm[k] = position of X[k] after dif_inplace."""
    N = int(np.prod(radices))
    out = np.zeros(N, dtype=np.int64)
    for k in range(N):
        kk, pos, div = k, 0, N
        for r in radices:
            div //= r
            pos += (kk % r) * div
            kk //= r
        out[k] = pos
    return out


def bluestein_inplace(z, M=None):
    """Bluestein with DIF forward / pointwise in permuted order / DIT inverse (no reordering pass)."""
    z = np.asarray(z, dtype=np.complex128)
    N = len(z)
    M = M or next_pow2(2 * N - 1)
    rad = factorize(M)
    j = np.arange(N)
    c = np.exp(-1j * np.pi * ((j * j) % (2 * N)) / N)
    b = np.zeros(M, np.complex128)
    b[:N] = np.conj(c)
    b[M - N + 1:] = np.conj(c[1:][::-1])
    bhat_perm = dif_inplace(b, rad) / M            # host-side table, stored in DIF order, 1/M folded in
    a = np.zeros(M, np.complex128)
    a[:N] = z * c
    A = dif_inplace(a, rad)
    C = np.conj(A * bhat_perm)
    D = dit_inplace(C, rad)
    return np.conj(D[:N]) * c
