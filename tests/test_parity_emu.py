"""Kernel-logic parity on the CPU: the CUDA kernel sources compiled with g++ against tests/emu/simt_emu.h
(one fiber per CUDA thread) and compared with the oracle.  Small sizes only; the parity tests proper are
tests/test_parity_gpu.py (-m gpu), which run the same cases through the nvcc-built library on a B200."""
import numpy as np
import pytest

from emu_backend import emu_backend
from parity_cases import Harness, seeded
from oracle import ndrustfft_oracle as orc

OPS = ["ndfft", "ndifft", "ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4"]


@pytest.fixture(scope="module")
def hs():
    return Harness(emu_backend())


def test_reference_unit_tests(hs):
    hs.reference_unit_tests()


# every radix, their products, Bluestein lengths, n = 1, 2
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 16, 17, 23, 26, 32, 45, 49, 60, 64, 77, 97, 121, 128, 169, 240])
@pytest.mark.parametrize("op", OPS)
def test_lengths_contiguous(hs, op, n):
    if op == "nddct1" and n < 2:
        pytest.skip("DCT-I needs n >= 2")
    hs.run(op, n, (3, n), 1, np.float64, seed=n)


@pytest.mark.parametrize("n", [2, 6, 12, 17, 30, 64, 100])
@pytest.mark.parametrize("op", OPS)
def test_f32(hs, op, n):
    hs.run(op, n, (4, n), 1, np.float32, seed=100 + n)
    hs.run(op, n, (n, 5), 0, np.float32, seed=200 + n)


@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("shape,axis", [((12, 7), 0), ((5, 12, 3), 1), ((12, 3, 4), 0), ((2, 3, 12), 2), ((2, 3, 2, 12, 2), 3), ((12,), 0)])
def test_layout_paths(hs, op, shape, axis):
    n = shape[axis]
    hs.run(op, n, shape, axis, np.float64, seed=7)          # path A / B
    hs.run(op, n, shape, axis, np.float64, seed=8, order="F")  # path C


@pytest.mark.parametrize("op", OPS)
def test_norm_none(hs, op):
    hs.run(op, 10, (3, 10), 1, np.float64, norm="none", seed=3)
    hs.run(op, 9, (9, 4), 0, np.float64, norm="none", seed=4)


@pytest.mark.parametrize("op", ["ndfft", "nddct2", "ndfft_r2c"])
def test_par_aliases(hs, op):
    hs.run(op, 12, (5, 12), 1, par=True)


def test_many_lanes_tail_tiles(hs):
    # lane counts that are not multiples of the tile width, both layouts
    hs.run("ndfft", 16, (37, 16), 1, seed=1)
    hs.run("ndfft", 16, (16, 37), 0, seed=2)
    hs.run("nddct2", 20, (20, 3, 11), 0, seed=3)


def test_strided_views_and_negative_strides(hs):
    be = hs.be
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, (8, 20)) + 1j * rng.uniform(-1, 1, (8, 20))
    x = base[::2, ::-2]                       # shape (4, 10), negative stride along the axis
    assert x.strides[1] < 0
    out_base = np.full((4, 25), 7 + 7j)
    y = out_base[:, 3:23:2]                    # non-dense output view: holes must be preserved
    h = be.FftHandler(10)
    be.ndfft(x, y, h, 1)
    yo = np.zeros((4, 10), complex)
    orc.ndfft(np.ascontiguousarray(x), yo, orc.FftHandler(10), 1)
    assert orc.rel_l2(y, yo) < 1e-12
    mask = np.ones(25, bool); mask[3:23:2] = False
    assert np.all(out_base[:, mask] == 7 + 7j)


def test_custom_normalization_all_kinds(hs):
    be = hs.be
    Norm = type(be.FftHandler(4).norm)
    rng = np.random.default_rng(11)

    def f(lane):
        lane *= 0.25
        lane[0] += 1.0

    n = 8
    # c2c inverse: after the transform (src/lib.rs:329)
    x = rng.uniform(-1, 1, (3, n)) + 1j * rng.uniform(-1, 1, (3, n))
    y = np.zeros_like(x); yo = np.zeros_like(x)
    be.ndifft(x, y, be.FftHandler(n).normalization(Norm.Custom(f)), 1)
    orc.ndifft(x, yo, orc.FftHandler(n).normalization(orc.Normalization.custom(f)), 1)
    assert orc.rel_l2(y, yo) < 1e-12
    # c2r: on the m-long spectrum copy, before; Im(DC)/Im(Nyq) zeroed afterwards (src/lib.rs:511-521)
    sp = rng.uniform(-1, 1, (3, n // 2 + 1)) + 1j * rng.uniform(-1, 1, (3, n // 2 + 1))
    r = np.zeros((3, n)); ro = np.zeros((3, n))
    be.ndifft_r2c(sp, r, be.R2cFftHandler(n).normalization(Norm.Custom(f)), 1)
    orc.ndifft_r2c(sp, ro, orc.R2cFftHandler(n).normalization(orc.Normalization.custom(f)), 1)
    assert orc.rel_l2(r, ro) < 1e-12
    # dct: on the input copy, before (src/lib.rs:692-696)
    xr = rng.uniform(-1, 1, (n, 3))
    for k in (1, 2, 3, 4):
        r = np.zeros((n, 3)); ro = np.zeros((n, 3))
        getattr(be, f"nddct{k}")(xr, r, be.DctHandler(n).normalization(Norm.Custom(f)), 0)
        getattr(orc, f"nddct{k}")(xr, ro, orc.DctHandler(n).normalization(orc.Normalization.custom(f)), 0)
        assert orc.rel_l2(r, ro) < 1e-12


def test_roundtrip_identities(hs):
    be = hs.be
    n = 24
    x = seeded(1, (5, n), np.float64, False)
    a = np.zeros_like(x); b = np.zeros_like(x)
    h = be.DctHandler(n)
    be.nddct2(x, a, h, 1); be.nddct3(a, b, h, 1)
    assert orc.rel_l2(b, 2 * n * x) < 1e-12
    be.nddct1(x, a, h, 1); be.nddct1(a, b, h, 1)
    assert orc.rel_l2(b, 2 * (n - 1) * x) < 1e-12
    be.nddct4(x, a, h, 1); be.nddct4(a, b, h, 1)
    assert orc.rel_l2(b, 2 * n * x) < 1e-12
    xc = seeded(2, (n, 3), np.float64, True)
    ya = np.zeros_like(xc); yb = np.zeros_like(xc)
    hf = be.FftHandler(n)
    be.ndfft(xc, ya, hf, 0); be.ndifft(ya, yb, hf, 0)
    assert orc.rel_l2(yb, xc) < 1e-12


def test_errors(hs):
    be = hs.be
    h = be.FftHandler(6)
    with pytest.raises(AssertionError, match="Size mismatch in fft, got 5 expected 6"):
        be.ndfft(np.zeros((2, 5), complex), np.zeros((2, 5), complex), h, 1)
    with pytest.raises(AssertionError, match="Size mismatch in fft, got 7 expected 6"):
        be.ndfft(np.zeros((2, 6), complex), np.zeros((2, 7), complex), h, 1)
    with pytest.raises(AssertionError, match="Size mismatch in dct, got 5 expected 6"):
        be.nddct3(np.zeros((5, 2)), np.zeros((5, 2)), be.DctHandler(6), 0)
    hr = be.R2cFftHandler(6)
    with pytest.raises(AssertionError, match="Size mismatch in fft, got 6 expected 4"):
        be.ndfft_r2c(np.zeros((2, 6)), np.zeros((2, 6), complex), hr, 1)
    with pytest.raises(AssertionError):   # ndarray Zip shape mismatch
        be.ndfft(np.zeros((2, 6), complex), np.zeros((3, 6), complex), h, 1)
    with pytest.raises(IndexError):
        be.ndfft(np.zeros((2, 6), complex), np.zeros((2, 6), complex), h, 2)
    with pytest.raises(TypeError):
        be.ndfft(np.zeros((2, 6)), np.zeros((2, 6), complex), h, 1)


def test_empty_arrays(hs):
    be = hs.be
    be.ndfft(np.zeros((0, 6), complex), np.zeros((0, 6), complex), be.FftHandler(6), 1)
    be.ndfft(np.zeros((3, 0), complex), np.zeros((3, 0), complex), be.FftHandler(0), 1)


def test_four_step_decomposition_small(hs):
    """The two-pass path (exec_four_step) is normally taken only for rows that overflow shared memory; the
    NDFB_FORCE_FOUR_STEP hook lets the emulator exercise it at a size it can finish."""
    import os
    os.environ["NDFB_FORCE_FOUR_STEP"] = "1"
    try:
        hs.run("ndfft", 64, (3, 64), 1, seed=5)
        hs.run("ndifft", 60, (2, 60), 1, seed=6)
        hs.run("ndfft", 48, (48, 3), 0, seed=7)
        hs.run("ndifft", 36, (2, 36, 2), 1, np.float32, seed=8)
    finally:
        del os.environ["NDFB_FORCE_FOUR_STEP"]


# ---- the instantiated Stockham fast path (sfft_kernel.cuh): every registered length, both layouts ----
SFFT_SIZES = [64, 128, 256, 512, 1024, 2048, 4096, 8192, 36, 60, 100, 216, 360, 384, 600, 1000]


@pytest.mark.parametrize("n", SFFT_SIZES)
def test_sfft_registered_lengths(hs, n, capfd):
    import os
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run("ndfft", n, (3, n), 1, np.float64, seed=n)
        hs.run("ndifft", n, (n, 5), 0, np.float64, seed=n + 1)
        hs.run("ndifft", n, (2, n), 1, np.float32, seed=n + 2)
        hs.run("ndfft", n, (n, 9), 0, np.float32, seed=n + 3)
        hs.run("ndfft", n, (2, n, 3), 1, np.float32, norm="none", seed=n + 4)
    finally:
        del os.environ["NDFB_TRACE"]
    err = capfd.readouterr().err
    # the fast path, not the general kernel, ran (f64 8192 strided columns have no schedule: 139 KB x 2 lanes)
    # (long f32 columns take the two-pass strided route: two fast-path launches per call)
    assert err.count("[ndfb] sfft") >= (4 if n == 8192 else 5), err


def test_sfft_and_general_kernel_agree(hs):
    """Same inputs through both kernel families (NDFB_DISABLE_SFFT is read once per process, so compare via oracle)."""
    hs.run("ndfft", 16384, (1, 16384), 1, np.float32, seed=1)     # f32 16384 has no schedule: general kernel
    hs.run("ndfft", 8192, (1, 8192), 1, np.float32, seed=1)       # fast path


def test_four_step_through_fast_path(hs):
    import os
    os.environ["NDFB_FORCE_FOUR_STEP"] = "1"
    try:
        hs.run("ndfft", 4096, (2, 4096), 1, np.float32, seed=5)      # 64 x 64, both passes on sfft schedules
        hs.run("ndifft", 8192, (2, 8192), 1, np.float64, seed=6)     # 64 x 128
        hs.run("ndfft", 4096, (4096, 3), 0, np.float64, seed=7)
    finally:
        del os.environ["NDFB_FORCE_FOUR_STEP"]


# ---- real-transform fast path (rsfft_kernel): R2C / C2R / DCT-I..IV around the Stockham core ----
@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("op", ["ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4"])
def test_rsfft_registered_lengths(hs, op, n, capfd):
    import os
    nn = n + 1 if op == "nddct1" else n      # DCT-I of n = 2^k + 1 points runs a 2^k-point core (benches/ndrustfft.rs:7)
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run(op, nn, (3, nn), 1, np.float64, seed=n)
        hs.run(op, nn, (nn, 5), 0, np.float32, seed=n + 1)
    finally:
        del os.environ["NDFB_TRACE"]
    err = capfd.readouterr().err
    # core length n/2 = 32 has a column schedule only; everything longer runs both layouts on the fast path
    core = n if op == "nddct1" else n // 2
    want = (1 if 64 <= core <= 4096 else 0) + (1 if 32 <= core <= 4096 else 0)   # registered row / column schedules
    assert err.count("[ndfb] rsfft") == want, err


def test_rsfft_unaligned_rows_fall_back(hs, capfd):
    """R2C rows read (x[2j], x[2j+1]) as one vector: an odd element offset must take the general kernel instead."""
    import os
    be = hs.be
    rng = np.random.default_rng(3)
    base = rng.uniform(-1, 1, (4, 129))
    x = base[:, 1:]                         # rows start on an odd element
    y = np.zeros((4, 65), complex); yo = np.zeros((4, 65), complex)
    os.environ["NDFB_TRACE"] = "1"
    try:
        # through the DEVICE entry (the emulation build's "device" memory is host memory): host views with gaps are packed
        # into an aligned staging copy by the library, so only device views can present an odd element offset
        import ctypes
        from ndrustfft_b200 import _lib
        h = be.R2cFftHandler(128)
        SZ, PD = ctypes.c_size_t * 2, ctypes.c_ssize_t * 2
        rc = be.lib.dll.ndfb_exec(h._plan, _lib.OP_R2C, _lib.NORM_DEFAULT, ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(y.ctypes.data), 2,
                                  SZ(4, 128), PD(129, 1), SZ(4, 65), PD(65, 1), 1, _lib.MEM_DEVICE, None)
        be.lib.check(rc)
    finally:
        del os.environ["NDFB_TRACE"]
    orc.ndfft_r2c(np.ascontiguousarray(x), yo, orc.R2cFftHandler(128), 1)
    assert orc.rel_l2(y, yo) < 1e-12
    assert "[ndfb] rsfft" not in capfd.readouterr().err


def test_rsfft_tail_tiles_and_3d(hs):
    hs.run("nddct2", 128, (37, 128), 1, seed=1)
    hs.run("nddct3", 128, (128, 37), 0, seed=2)
    hs.run("ndfft_r2c", 256, (3, 256, 5), 1, seed=3)
    hs.run("ndifft_r2c", 256, (2, 3, 256), 2, np.float32, seed=4, norm="none")
    hs.run("nddct4", 64, (64, 3, 7), 0, np.float32, seed=5)


def test_strided_four_step(hs, capfd):
    """Long strided columns whose one-pass tile would be narrower than a 32-byte sector take the two-pass route with the
    adjacent columns kept innermost in the workspace (c2 axis 0 in BASELINE.json)."""
    import os
    os.environ["NDFB_TRACE"] = "1"
    os.environ["NDFB_STRIDED_FOURSTEP"] = "1"
    try:
        hs.run("ndfft", 1024, (1024, 20), 0, np.float32, seed=1)
        hs.run("ndifft", 2048, (2048, 7), 0, np.float64, seed=2)
        hs.run("ndfft", 1024, (3, 1024, 5), 1, np.float32, seed=3)
    finally:
        del os.environ["NDFB_TRACE"]
        del os.environ["NDFB_STRIDED_FOURSTEP"]
    err = capfd.readouterr().err
    assert err.count("(strided lanes)") == 3, err


def test_split_output_axis(hs):
    """ndfb_exec_split_out: the packed all-to-all send layout written directly by the axis pass."""
    be = hs.be
    rng = np.random.default_rng(9)
    s0, n1, mc, P = 3, 12, 5, 4
    s1 = n1 // P
    x = rng.uniform(-1, 1, (s0, n1, mc)) + 1j * rng.uniform(-1, 1, (s0, n1, mc))
    send = np.zeros(P * s0 * s1 * mc, complex)
    be.ndfft_split_out(x, send, be.FftHandler(n1), 1, out_shape=(s0, n1, mc), out_strides=(s1 * mc, mc, 1),
                       out_block=s1, out_block_stride=s0 * s1 * mc)
    want = np.fft.fft(x, axis=1).reshape(s0, P, s1, mc).transpose(1, 0, 2, 3).ravel()
    assert orc.rel_l2(send, want) < 1e-12
    # 64-point lanes take the Stockham fast path; same layout contract
    n1, P = 64, 2
    s1 = n1 // P
    x = rng.uniform(-1, 1, (s0, n1, mc)) + 1j * rng.uniform(-1, 1, (s0, n1, mc))
    send = np.zeros(P * s0 * s1 * mc, complex)
    be.ndfft_split_out(x, send, be.FftHandler(n1), 1, out_shape=(s0, n1, mc), out_strides=(s1 * mc, mc, 1),
                       out_block=s1, out_block_stride=s0 * s1 * mc, inverse=True)
    want = np.fft.ifft(x, axis=1).reshape(s0, P, s1, mc).transpose(1, 0, 2, 3).ravel()
    assert orc.rel_l2(send, want) < 1e-12


def test_dct1_4096_runs_the_4095_point_schedule(hs, capfd):
    """BASELINE c4: nddct1 on 4096 points needs a 4095 = 13 * 9 * 7 * 5 point core (radix-13/9/7/5 Stockham passes)."""
    import os
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run("nddct1", 4096, (2, 4096), 1, np.float64, seed=1)
        hs.run("ndfft", 729, (729, 3), 0, np.float32, seed=2)       # 9 * 9 * 9
        hs.run("ndifft", 81, (4, 81), 1, np.float64, seed=3)
    finally:
        del os.environ["NDFB_TRACE"]
    err = capfd.readouterr().err
    assert "rsfft kind=2 f64 N=4095" in err and err.count("[ndfb] sfft") == 2, err


def test_scatter_output_blocks(hs):
    """ndfb_exec_scatter_out: each destination block of the output lanes goes to its own buffer (peer receive buffers on
    a multi-GPU box; here: separate host arrays under the emulator)."""
    be = hs.be
    rng = np.random.default_rng(10)
    s0, n1, mc, P = 3, 128, 5, 4
    s1 = n1 // P
    x = rng.uniform(-1, 1, (s0, n1, mc)) + 1j * rng.uniform(-1, 1, (s0, n1, mc))
    bufs = [np.zeros((s0, s1, mc), complex) for _ in range(P)]
    be.ndfft_scatter_out(x, be.FftHandler(n1), 1, out_shape=(s0, n1, mc), out_strides=(s1 * mc, mc, 1), out_block=s1,
                         block_ptrs=[b.ctypes.data for b in bufs])
    want = np.fft.fft(x, axis=1)
    for p in range(P):
        assert orc.rel_l2(bufs[p], want[:, p * s1:(p + 1) * s1, :]) < 1e-12


# ---- staged path (big_kernels.cuh): prologue kernel -> complex core on workspace rows -> epilogue kernel ----
@pytest.mark.parametrize("op", ["ndfft", "ndifft", "ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4"])
@pytest.mark.parametrize("n", [64, 90, 202, 1018, 45, 101, 1009])     # even: smooth, smooth, 2*101 (Bluestein core), 2*509; odd: smooth, prime, prime
def test_staged_path(hs, op, n, capfd):
    """Lengths that overflow one CTA's shared memory take the staged route; NDFB_FORCE_STAGED runs it at test sizes."""
    import os
    nn = n + 1 if op == "nddct1" else n
    os.environ["NDFB_FORCE_STAGED"] = "1"
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run(op, nn, (3, nn), 1, np.float64, seed=n)
        hs.run(op, nn, (nn, 5), 0, np.float32, seed=n + 1, norm="none")
    finally:
        del os.environ["NDFB_FORCE_STAGED"]
        del os.environ["NDFB_TRACE"]
    assert capfd.readouterr().err.count("[ndfb] staged") == 2


def test_staged_path_prime_c2c(hs):
    import os
    os.environ["NDFB_FORCE_STAGED"] = "1"
    try:
        hs.run("ndfft", 1009, (2, 1009), 1, np.float64, seed=1)
        hs.run("ndifft", 257, (257, 3), 0, np.float64, seed=2)
        hs.run("ndfft", 97, (4, 97, 2), 1, np.float32, seed=3)
    finally:
        del os.environ["NDFB_FORCE_STAGED"]


def test_three_pass_decomposition(hs, capfd):
    """Rows too long for two on-chip factors split the second factor again (2^24 = 256 x (256 x 256) on the GPU);
    NDFB_FS_CAP shrinks the 'chip' so the emulator can run the same code at 16384 = 16 x (32 x 32)."""
    import os
    os.environ.update({"NDFB_FORCE_FOUR_STEP": "1", "NDFB_FS_CAP": "64", "NDFB_FS_N1": "16", "NDFB_TRACE": "1"})
    try:
        hs.run("ndfft", 16384, (2, 16384), 1, np.float32, seed=1)
        hs.run("ndifft", 16384, (1, 16384), 1, np.float64, seed=2)
    finally:
        for k in ("NDFB_FORCE_FOUR_STEP", "NDFB_FS_CAP", "NDFB_FS_N1", "NDFB_TRACE"):
            del os.environ[k]
    assert capfd.readouterr().err.count("second factor split again") == 2


# ---- fused Bluestein on the Stockham passes (bsfft_kernel): lengths with no schedule of their own ----
@pytest.mark.parametrize("n", [17, 31, 97, 257, 1009, 2047, 2049])
def test_bsfft_lengths(hs, n, capfd):
    import os
    os.environ["NDFB_TRACE"] = "1"
    try:
        if n <= 2048:                                    # f64: M = 4096 is the largest instantiated length
            hs.run("ndfft", n, (3, n), 1, np.float64, seed=n)
            hs.run("ndifft", n, (n, 5), 0, np.float64, seed=n + 1)
        hs.run("ndifft", n, (2, n), 1, np.float32, seed=n + 2)
        hs.run("ndfft", n, (n, 9), 0, np.float32, seed=n + 3)
        hs.run("ndfft", n, (2, n, 3), 1, np.float32, norm="none", seed=n + 4)
    finally:
        del os.environ["NDFB_TRACE"]
    # (f64 columns at M = 4096 and f32 columns at M = 8192 have no tile that fits: those calls stay on the general kernel)
    assert capfd.readouterr().err.count("[ndfb] bsfft") == (5 if n <= 1024 else 4 if n <= 2048 else 1)


def test_bsfft_and_general_bluestein_agree(hs, capfd):
    import os
    os.environ["NDFB_DISABLE_BSFFT"] = "1"
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run("ndfft", 1009, (2, 1009), 1, np.float64, seed=5)
        hs.run("ndfft", 97, (97, 4), 0, np.float32, seed=6)
    finally:
        del os.environ["NDFB_DISABLE_BSFFT"]
        del os.environ["NDFB_TRACE"]
    assert "[ndfb] bsfft" not in capfd.readouterr().err


# ---- multi-axis chains (ndfb_exec_chain): intermediates in `out` or in library workspaces ----
def test_chain_reference_examples(hs):
    hs.chain_examples()


@pytest.mark.parametrize("rd", [np.float64, np.float32])
def test_chain_fft_rfft_nd(hs, rd):
    hs.run_chain([("ndfft", 12, 1), ("ndfft", 10, 0)], (10, 12), rd, seed=1)                              # fft2: second step in place on out
    hs.run_chain([("ndifft", 10, 0), ("ndifft", 12, 1)], (10, 12), rd, seed=2)
    hs.run_chain([("ndfft_r2c", 16, 2), ("ndfft", 6, 1), ("ndfft", 5, 0)], (5, 6, 16), rd, seed=3)       # rfft3 (config 3 pattern)
    hs.run_chain([("ndifft", 5, 0), ("ndifft", 6, 1), ("ndifft_r2c", 16, 2)], (5, 6, 9), rd, seed=4)     # irfft3: workspace, in place on it
    hs.run_chain([("ndifft", 5, 0), ("ndifft_r2c", 9, 1)], (5, 5), rd, seed=5, norm="none")              # odd real length
    hs.run_chain([("ndfft", 64, 0)], (64, 3), rd, seed=6)                                                  # single step = plain call


def test_chain_dct_and_mixed(hs):
    hs.run_chain([("nddct2", 8, 0), ("nddct3", 8, 0)], (8, 5), seed=1)                                   # 2n * x
    hs.run_chain([("nddct1", 9, 1), ("nddct4", 6, 0), ("nddct2", 4, 2)], (6, 9, 4), seed=2)
    hs.run_chain([("nddct2", 10, 0), ("ndfft_r2c", 12, 1)], (10, 12), seed=3)                             # Chebyshev x Fourier
    hs.run_chain([("ndfft_r2c", 12, 1), ("ndifft_r2c", 12, 1)], (4, 12), seed=4)                          # real -> complex -> real
    hs.run_chain([("ndfft_r2c", 8, 1), ("ndfft", 6, 0), ("ndifft", 6, 0), ("ndifft_r2c", 8, 1)], (6, 8), seed=5)


def test_chain_in_place(hs):
    hs.run_chain([("ndfft", 12, 1), ("ndfft", 10, 0)], (10, 12), seed=1, inplace=True)
    hs.run_chain([("nddct2", 16, 0)], (16, 4), seed=2, inplace=True)
    hs.run("ndfft", 64, (64, 3), 0)  # (single calls in place are covered below)


@pytest.mark.parametrize("op,n,shape,axis", [("ndfft", 360, (3, 360), 1), ("ndifft", 64, (64, 5), 0), ("nddct1", 9, (9, 3), 0),
                                              ("nddct4", 12, (2, 12), 1), ("ndfft", 1009, (2, 1009), 1), ("ndfft", 17, (17, 4), 0)])
def test_single_call_in_place(hs, op, n, shape, axis):
    import oracle.ndrustfft_oracle as orc
    from parity_cases import seeded
    icx = hs.OPS[op][1]
    x = seeded(3, shape, np.float64, icx)
    h = getattr(hs.be, hs.OPS[op][0])(n)
    ho = getattr(orc, hs.OPS[op][0])(n)
    want = np.zeros(shape, np.complex128 if icx else np.float64)
    getattr(orc, op)(x, want, ho, axis)
    buf = hs.mk(x)
    getattr(hs.be, op)(buf, buf, h, axis)
    assert orc.rel_l2(hs.to_np(buf), want) <= 1e-12


def test_chain_in_place_four_step(hs):
    import os
    os.environ["NDFB_FORCE_FOUR_STEP"] = "1"
    try:
        hs.run_chain([("ndfft", 64, 1), ("ndfft", 36, 0)], (36, 64), seed=1, inplace=True)
    finally:
        del os.environ["NDFB_FORCE_FOUR_STEP"]


def test_chain_errors(hs):
    from ndrustfft_b200 import SizeMismatch, NdfftError
    be = hs.be
    x = np.zeros((4, 6), np.complex128); y = np.zeros((4, 6), np.complex128)
    with pytest.raises(SizeMismatch, match="Size mismatch in fft, got 6 expected 5"):
        be.ndchain(x, y, [("ndfft", be.FftHandler(5), 1), ("ndfft", be.FftHandler(4), 0)])
    with pytest.raises(SizeMismatch, match="Size mismatch in fft, got 4 expected 3"):
        be.ndchain(x, y, [("ndfft", be.FftHandler(6), 1), ("ndfft", be.FftHandler(3), 0)])
    with pytest.raises(IndexError):
        be.ndchain(x, y, [("ndfft", be.FftHandler(6), 2)])
    xr = np.zeros((4, 6)); yc = np.zeros((4, 6), np.complex128)
    with pytest.raises(SizeMismatch, match="Size mismatch in fft, got 6 expected 4"):
        be.ndchain(xr, yc, [("ndfft_r2c", be.R2cFftHandler(6), 1)])
    with pytest.raises(NdfftError, match="reads real data"):
        be.ndchain(x, np.zeros((4, 6)), [("ndfft", be.FftHandler(6), 1), ("nddct2", be.DctHandler(4), 0)])
    with pytest.raises(NdfftError, match="share dtype"):
        be.ndchain(x, y, [("ndfft", be.FftHandler(6), 1), ("ndfft", be.FftHandler(4, np.float32), 0)])
    with pytest.raises(ValueError):
        be.ndchain(x, y, [])


def test_chain_custom_normalisation_falls_back_to_single_calls(hs):
    be = hs.be
    Norm = type(be.FftHandler(3).norm)

    def my_norm(lane):
        lane *= 0.5

    rng = np.random.default_rng(0)
    x = (rng.uniform(-1, 1, (6, 8)) + 1j * rng.uniform(-1, 1, (6, 8)))
    y = np.zeros_like(x); want = np.zeros_like(x); work = np.zeros_like(x)
    h0, h1 = be.FftHandler(6).normalization(Norm.Custom(my_norm)), be.FftHandler(8)
    be.ndchain(x, y, [("ndifft", h0, 0), ("ndifft", h1, 1)])
    be.ndifft(x, work, h0, 0); be.ndifft(work, want, h1, 1)
    assert np.allclose(y, want, rtol=0, atol=1e-14)


def test_four_step_l2_groups(hs, capfd):
    """Two-pass transforms run group by group over a reused workspace (L2-resident on the GPU); NDFB_FS_L2_KB shrinks the
    group size so the emulator exercises the slicing of the innermost column dim, of outer batch dims and of both."""
    import os
    os.environ.update({"NDFB_FORCE_FOUR_STEP": "1", "NDFB_FS_L2_KB": "64", "NDFB_TRACE": "1"})
    try:
        hs.run("ndfft", 64, (64, 200), 0, np.float32, seed=1)            # strided columns: 64 x 200 x 8 B = 100 KB -> column groups of 128
        n1 = capfd.readouterr().err.count("[ndfb] four-step")
        hs.run("ndifft", 256, (12, 256, 3), 1, np.float64, seed=2)       # outer dim sliced (5 + 5 + 2), 3 columns kept together
        n2 = capfd.readouterr().err.count("[ndfb] four-step")
        hs.run("ndfft", 1024, (20, 1024), 1, np.float32, seed=3)         # contiguous rows: 8 KB each -> groups of 8 rows
        n3 = capfd.readouterr().err.count("[ndfb] four-step")
        hs.run("ndfft", 64, (3, 64, 130), 1, np.float64, seed=4)         # outer dim to single indices, then column groups
        n4 = capfd.readouterr().err.count("[ndfb] four-step")
        hs.run("ndfft", 16384, (2, 16384), 1, np.float64, seed=5)        # one lane is bigger than a group: not sliced
        n5 = capfd.readouterr().err.count("[ndfb] four-step")
    finally:
        for k in ("NDFB_FORCE_FOUR_STEP", "NDFB_FS_L2_KB", "NDFB_TRACE"):
            del os.environ[k]
    assert (n1, n2, n3, n4, n5) == (2, 3, 3, 9, 1), (n1, n2, n3, n4, n5)


def test_fused_two_pass_columns(hs, capfd):
    """Long strided columns: both four-step passes in one persistent launch over an L2-sized workspace ring (fs2_kernel).
    NDFB_FS2_KB shrinks the group budget so that small arrays already form >= 4 groups."""
    import os
    os.environ.update({"NDFB_TRACE": "1", "NDFB_FS2": "1", "NDFB_FS2_KB": "4096", "NDFB_FS2_F64": "1", "NDFB_FS2_ALL": "1"})
    try:
        hs.run("ndfft", 8192, (8192, 256), 0, np.float32, seed=1)                    # 64 x 128, 4 groups of 64 columns
        hs.run("ndifft", 8192, (8192, 192), 0, np.float64, seed=2, norm="none")      # 6 groups of 32 columns: ring slots reused
        os.environ["NDFB_FS2_KB"] = "8192"
        hs.run("ndifft", 16384, (16384, 256), 0, np.float32, seed=3)                 # 128 x 128, 4 groups of 64
        err = capfd.readouterr().err
        assert err.count("[ndfb] fs2") == 3 and "four-step" not in err, err
        hs.run("ndfft", 8192, (8192, 100), 0, np.float32, seed=4)                    # 100 columns: no group width divides -> two launches
        os.environ["NDFB_NO_FS2"] = "1"
        hs.run("ndfft", 8192, (8192, 256), 0, np.float32, seed=1)
        err = capfd.readouterr().err
        assert "[ndfb] fs2" not in err and err.count("[ndfb] four-step") == 2
    finally:
        for k in ("NDFB_TRACE", "NDFB_FS2", "NDFB_FS2_KB", "NDFB_NO_FS2", "NDFB_FS2_F64", "NDFB_FS2_ALL"):
            os.environ.pop(k, None)


def test_dct1_4096_strided_axis_runs_the_4095_point_tile(hs, capfd):
    import os
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run("nddct1", 4096, (4096, 5), 0, np.float64, seed=1)
        hs.run("nddct1", 4096, (4096, 3), 0, np.float32, seed=2, norm="none")
    finally:
        del os.environ["NDFB_TRACE"]
    err = capfd.readouterr().err
    assert err.count("[ndfb] rsfft kind=2") == 2 and "N=4095 cols L=2" in err, err


def test_schedule_selection_for_the_baseline_shapes(hs, capfd):
    """The host's tile / variant choice for the BASELINE shapes (selection rules in find_sfft / find_rsfft; measured A/B
    behind each rule in profiles/): regression guard, run on scaled-down batches (the rules look at strides, not at counts)."""
    import os
    os.environ["NDFB_TRACE"] = "1"
    be = hs.be
    try:
        def trace(fn, h, shape_in, shape_out, axis, cplx_in=True, cplx_out=True, rd=np.float64):
            cd_ = np.complex128 if rd == np.float64 else np.complex64
            x = np.zeros(shape_in, cd_ if cplx_in else rd); y = np.zeros(shape_out, cd_ if cplx_out else rd)
            capfd.readouterr()
            getattr(be, fn)(x, y, h, axis)
            return capfd.readouterr().err
        # c5a: mixed-radix c128, register-capped variants (2 CTAs/SM) on every axis
        e = trace("ndfft", be.FftHandler(360), (360, 8, 384), (360, 8, 384), 0)
        assert "N=360 cols L=8 T=480" in e and "minb=2" in e, e
        e = trace("ndfft", be.FftHandler(1000), (2, 1000, 384), (2, 1000, 384), 1)
        assert "N=1000 cols L=4 T=400" in e and "minb=2" in e, e
        e = trace("ndfft", be.FftHandler(384), (2, 8, 384), (2, 8, 384), 2)
        assert "N=384 rows L=4 T=256" in e and "minb=4" in e, e
        # c3: 512-point c128 columns take the 8-lane family-B tile; r2c rows the 256-point core
        e = trace("ndfft", be.FftHandler(512), (512, 16, 257), (512, 16, 257), 0)
        assert "N=512 cols L=8 T=512" in e and "fam=B" in e, e
        e = trace("ndfft_r2c", be.R2cFftHandler(512), (4, 4, 512), (4, 4, 257), 2, cplx_in=False)
        assert "rsfft kind=0 f64 N=256 rows L=8 T=256" in e, e
        # c4: DCT-I of 4096 points = 4095-point core as 15.13.7.3 on 320 threads (two CTAs per SM in f64) and the two-column strided tile
        e = trace("nddct1", be.DctHandler(4096), (2, 4096), (2, 4096), 1, cplx_in=False, cplx_out=False)
        assert "N=4095 rows L=1 T=320" in e and "minb=2" in e, e
        e = trace("nddct1", be.DctHandler(4096), (4096, 4), (4096, 4), 0, cplx_in=False, cplx_out=False)
        assert "N=4095 cols L=2 T=640" in e, e
        # c2: 8192-point c64 rows = family B 16.16.16.2 with 512 threads; strided columns = two passes 64 x 128
        e = trace("ndfft", be.FftHandler(8192, np.float32), (2, 8192), (2, 8192), 1, rd=np.float32)
        assert "N=8192 rows L=1 T=512" in e and "fam=B" in e and "minb=2" in e, e
        e = trace("ndfft", be.FftHandler(8192, np.float32), (8192, 64), (8192, 64), 0, rd=np.float32)
        assert "four-step N=8192 = 64 x 128 (strided lanes)" in e and "fs2" not in e, e
    finally:
        del os.environ["NDFB_TRACE"]


def test_three_pass_transposing_last_pass(hs, capfd):
    """Last pass of a three-pass split as the transposing rows kernel (contiguous rows in, lane-interleaved out):
    64^3 = 64 x (64 x 64) with the 'chip' shrunk by NDFB_FS_CAP; both precisions; and the capped column tile as fallback."""
    import os
    os.environ.update({"NDFB_FORCE_FOUR_STEP": "1", "NDFB_FS_CAP": "256", "NDFB_FS_N1": "64", "NDFB_TRACE": "1"})
    try:
        hs.run("ndfft", 64 ** 3, (2, 64 ** 3), 1, np.float32, seed=1)
        hs.run("ndifft", 64 ** 3, (1, 64 ** 3), 1, np.float64, seed=2)
        err = capfd.readouterr().err
        assert err.count("rows->lanes (transposing)") == 2, err
        os.environ["NDFB_NO_TRANS_STORE"] = "1"
        hs.run("ndfft", 64 ** 3, (1, 64 ** 3), 1, np.float32, seed=3)
        assert "transposing" not in capfd.readouterr().err
    finally:
        for k in ("NDFB_FORCE_FOUR_STEP", "NDFB_FS_CAP", "NDFB_FS_N1", "NDFB_TRACE", "NDFB_NO_TRANS_STORE"):
            os.environ.pop(k, None)


def test_rows_bulk_async_kernel(hs, capfd):
    """Persistent bulk-async row kernel (TMA loads / stores + mbarrier on the GPU; plain copies under the emulator):
    several tiles per CTA, a ragged last tile, forward and inverse, both precisions."""
    import os
    os.environ.update({"NDFB_ROWS_BULK": "2", "NDFB_TRACE": "1"})
    try:
        hs.run("ndfft", 1024, (11, 1024), 1, np.float32, seed=1)            # L = 4: 3 tiles, the last one ragged
        hs.run("ndifft", 512, (9, 512), 1, np.float64, seed=2)
        hs.run("ndfft", 512, (3, 5, 512), 2, np.float64, seed=3, norm="none")
        err = capfd.readouterr().err
        assert err.count("rows bulk-async persistent") == 3, err
    finally:
        os.environ.pop("NDFB_ROWS_BULK", None); os.environ.pop("NDFB_TRACE", None)


@pytest.mark.parametrize("op,n,rd", [("ndfft_r2c", 512, np.float64), ("nddct2", 4096, np.float64), ("nddct1", 2049, np.float64),
                                     ("ndfft_r2c", 4096, np.float32), ("nddct2", 512, np.float64), ("nddct2", 4096, np.float32)])
def test_mirror_paired_last_pass(hs, op, n, rd):
    """Schedules whose last pass has two butterflies per thread run the pair epilogue from registers (thread i owns
    butterflies i and NB - i: bins k and N - k): 256-point f64 core (8.8.4: c3's r2c), 2048-point f64 core (8.8.8.4: c4's
    DCT-II / DCT-I of 2049), 2048-point f32 core (16.16.8); rows and columns, ragged tiles."""
    hs.run(op, n, (3, n), 1, rd, seed=n)
    hs.run(op, n, (n, 5), 0, rd, seed=n + 1, norm="none")


@pytest.mark.parametrize("op,n,rd", [("ndifft_r2c", 512, np.float64), ("nddct3", 4096, np.float64), ("ndifft_r2c", 4096, np.float32),
                                     ("nddct3", 512, np.float64), ("nddct3", 4096, np.float32)])
def test_mirror_paired_first_pass(hs, op, n, rd, capfd):
    """C2R / DCT-III on the small-radix-first schedules (4.8.8 / 4.8.8.8 / 8.16.16): pass 0 runs butterflies i and NB - i in one
    thread, fed by the zip of bins j and N - j straight from global memory (no prologue round trip through shared memory)."""
    import os
    os.environ["NDFB_TRACE"] = "1"
    os.environ["NDFB_MIRROR_PRO"] = "1"          # opt-in: measured slower than the big-radix-first schedules on B200
    try:
        hs.run(op, n, (3, n), 1, rd, seed=n)
        hs.run(op, n, (n, 5), 0, rd, seed=n + 1, norm="none")
    finally:
        del os.environ["NDFB_TRACE"]
        del os.environ["NDFB_MIRROR_PRO"]
    assert capfd.readouterr().err.count("fam=R") == 2


@pytest.mark.parametrize("op,n,rd", [("ndfft_r2c", 1024, np.float64), ("nddct2", 1024, np.float64), ("nddct1", 513, np.float64),
                                     ("ndfft_r2c", 8192, np.float32), ("nddct2", 2048, np.float64), ("ndfft_r2c", 2048, np.float32),
                                     ("nddct2", 1024, np.float32), ("ndfft_r2c", 256, np.float64)])
def test_pair_epilogue_from_registers_all_schedules(hs, op, n, rd):
    """One butterfly per thread in the last pass (512-point f64 core 8.8.8, 4096-point f32 core 16.16.16): mirror outputs
    arrive by warp shuffle (rows).  Four or eight butterflies per thread (1024-point f64 core 8.8.8.2, f32 cores 16.16.4 /
    16.16.2): G/2 in-thread mirror pairs.  Columns of the one-butterfly schedules keep the shared-memory epilogue."""
    hs.run(op, n, (3, n), 1, rd, seed=n)
    hs.run(op, n, (n, 5), 0, rd, seed=n + 1, norm="none")


def test_pipelined_column_kernel(hs, capfd):
    """Software-pipelined persistent column kernel (pipe_kernel.cuh: cp.async staging of the next tile, split re / im exchange;
    plain copies under the emulator): lane-adjacent input (strided axis) and row input (contiguous axis written lane-interleaved),
    several tiles per CTA, forward / inverse, full and partial passes (360 = 6.6.10), both precisions."""
    import os
    os.environ.update({"NDFB_PIPE": "2", "NDFB_TRACE": "1"})
    try:
        hs.run("ndfft", 64, (64, 12), 0, np.float32, seed=1)                 # INMODE 0: 3 tiles of 4 lanes
        hs.run("ndifft", 512, (512, 8), 0, np.float32, seed=2)               # 16.16.2
        hs.run("ndfft", 256, (2, 256, 8), 1, np.float32, seed=3, norm="none")
        hs.run("ndifft", 64, (64, 6), 0, np.float64, seed=4)                 # f64 8.8, L = 2
        hs.run("ndfft", 512, (512, 4), 0, np.float64, seed=5)
        hs.run("ndfft", 360, (360, 4), 0, np.float64, seed=6)                # partial passes
        err = capfd.readouterr().err
        assert err.count("cols pipelined") == 6, err
        assert err.count("in=lane-adjacent") == 6, err
        hs.run("ndfft", 256, (7, 256), 1, np.float32, seed=9)                # one-lane tiles: contiguous rows in and out
        hs.run("ndifft", 512, (3, 5, 512), 2, np.float64, seed=10)
        err = capfd.readouterr().err
        assert err.count("rows pipelined") == 2 and err.count("L=1 ") == 2, err
    finally:
        os.environ.pop("NDFB_PIPE", None); os.environ.pop("NDFB_TRACE", None)


def test_pipelined_column_kernel_four_step(hs, capfd):
    """Both passes of a two-pass split on the pipelined kernel: pass 1 reads lane-adjacent rows and applies the four-step twiddle
    in its store, pass 2 reads contiguous workspace rows (row staging layout) and stores lane-interleaved."""
    import os
    os.environ.update({"NDFB_PIPE": "2", "NDFB_TRACE": "1", "NDFB_FS_CAP": "512", "NDFB_FS_N1": "64", "NDFB_NO_FS_TRANSPOSE": "1"})
    try:
        hs.run("ndfft", 64 * 512, (2, 64 * 512), 1, np.float32, seed=7)
        hs.run("ndifft", 64 * 512, (1, 64 * 512), 1, np.float64, seed=8)
        err = capfd.readouterr().err
        assert err.count("in=lane-adjacent") == 2 and err.count("in=rows") == 2, err
        # beyond 2^17 points the four-step twiddle base is the hi/lo table product (times the per-lane factor table)
        os.environ["NDFB_FS_N1"] = "512"
        hs.run("ndfft", 512 * 512, (1, 512 * 512), 1, np.float32, seed=9)
        err = capfd.readouterr().err
        assert err.count("in=lane-adjacent") == 1 and err.count("in=rows") == 1, err
    finally:
        for k in ("NDFB_PIPE", "NDFB_TRACE", "NDFB_FS_CAP", "NDFB_FS_N1", "NDFB_NO_FS_TRANSPOSE"):
            os.environ.pop(k, None)


@pytest.mark.parametrize("op,n,rd", [("nddct3", 4096, np.float64), ("nddct4", 4096, np.float64), ("nddct3", 2048, np.float64), ("nddct4", 2048, np.float64),
                                     ("nddct3", 4096, np.float32), ("nddct4", 4096, np.float32), ("nddct4", 512, np.float64), ("nddct3", 1024, np.float32)])
def test_mirror_paired_output_pass(hs, op, n, rd, capfd):
    """DCT-III / DCT-IV on contiguous aligned rows: schedules whose last pass has an even number of butterflies per thread (8.8.8.4,
    8.8.8.2, 16.16.8, 8.8.4, 16.16.2) pair butterflies p and NB-1-p and store pairs of reals straight from registers.  Ragged tiles,
    default and no normalisation.  (Rows that start on an odd element keep the staged copy-out: GPU test, device views.)"""
    import os
    os.environ["NDFB_TRACE"] = "1"
    os.environ["NDFB_MIRROR_OUT"] = "1"      # (no longer needed: every DCT-III / DCT-IV row call takes it)
    try:
        hs.run(op, n, (3, n), 1, rd, seed=n)
        assert "mirror-paired output" in capfd.readouterr().err
        hs.run(op, n, (5, n), 1, rd, seed=n + 1, norm="none")
    finally:
        del os.environ["NDFB_TRACE"]; del os.environ["NDFB_MIRROR_OUT"]


def test_four_step_twiddle_factored(hs, capfd):
    """Column passes of the two- and three-pass splits form W_N^{k j2} as (tile-uniform lookup) x ([k][l] table): single-table
    lengths (N <= 2^17), the hi/lo product beyond, the nested split, forward and inverse; NDFB_NO_FS_FACTORED keeps the per-point
    lookups and gives the same result within rounding."""
    import os
    os.environ.update({"NDFB_TRACE": "1", "NDFB_FS_CAP": "512"})
    try:
        hs.run("ndfft", 64 * 512, (2, 64 * 512), 1, np.float32, seed=1)
        hs.run("ndifft", 128 * 256, (1, 128 * 256), 1, np.float64, seed=2)
        os.environ["NDFB_FS_N1"] = "512"
        hs.run("ndfft", 512 * 512, (1, 512 * 512), 1, np.float32, seed=3)          # 2^18: hi / lo tables
        del os.environ["NDFB_FS_N1"]
        err = capfd.readouterr().err
        assert err.count("fs twiddle factored") == 3, err
        os.environ.update({"NDFB_FS_CAP": "256", "NDFB_FS_N1": "64"})
        hs.run("ndfft", 64 * 64 * 64, (1, 64 * 64 * 64), 1, np.float64, seed=4)    # three passes: both column passes carry a twiddle
        err = capfd.readouterr().err
        assert err.count("fs twiddle factored") == 2, err
        del os.environ["NDFB_FS_N1"]
        os.environ["NDFB_NO_FS_FACTORED"] = "1"
        hs.run("ndfft", 64 * 64, (3, 64 * 64), 1, np.float32, seed=5)
        assert "fs twiddle factored" not in capfd.readouterr().err
    finally:
        for k in ("NDFB_TRACE", "NDFB_FS_CAP", "NDFB_FS_N1", "NDFB_NO_FS_FACTORED"):
            os.environ.pop(k, None)


def test_default_split_rules_at_real_sizes(hs, capfd):
    """The multi-pass split the host picks WITHOUT the test caps: f32 2^19 = 1024 x 512 in two passes with the
    transposing rows kernel as last pass; f64 2^19 takes three passes (64-point first factor) rather than a 1024-point one."""
    import os
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run("ndfft", 1 << 19, (1, 1 << 19), 1, np.float32, seed=1)
        err = capfd.readouterr().err
        assert "four-step N=524288 = 1024 x 512 (contiguous lanes)" in err and "N=512 rows->lanes (transposing)" in err, err
        hs.run("ndfft", 1 << 19, (1, 1 << 19), 1, np.float64, seed=2)
        err = capfd.readouterr().err
        assert "four-step N=524288 = 64 x 8192 (contiguous lanes, second factor split again)" in err, err
        assert err.count("fs twiddle factored") == 2 and "rows->lanes (transposing)" in err, err
        hs.run("ndifft", 1 << 21, (1, 1 << 21), 1, np.float32, seed=3)      # f32 from 2^21: 256 x (a x b)
        err = capfd.readouterr().err
        assert "four-step N=2097152 = 256 x 8192 (contiguous lanes, second factor split again)" in err, err
    finally:
        del os.environ["NDFB_TRACE"]


@pytest.mark.parametrize("op", ["ndfft_r2c", "ndifft_r2c", "nddct2", "nddct3", "nddct4"])
def test_real_kinds_on_the_4095_point_core(hs, op, capfd):
    """n = 8190 = 2 x 4095: every even-length real kind runs the 4095-point core, i.e. the measured schedule pick 15.13.7.3 on 320
    threads (unpadded: odd first radix), rows and columns, both precisions."""
    import os
    os.environ["NDFB_TRACE"] = "1"
    try:
        hs.run(op, 8190, (2, 8190), 1, np.float64, seed=3)
        hs.run(op, 8190, (8190, 3), 0, np.float32, seed=4)
    finally:
        del os.environ["NDFB_TRACE"]
    err = capfd.readouterr().err
    assert "N=4095 rows L=1 T=320" in err and "N=4095 cols L=2 T=640" in err, err
