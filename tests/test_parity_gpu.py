"""The parity tests proper: the nvcc-built sm_100a library, called through the C ABI, against the oracle.

Small/medium cases compare whole arrays; at BASELINE.json's full sizes (c2..c5) a seeded subset of lanes is
compared with the oracle (lanes are independent, src/lib.rs:120-124) and size-independent identities
(round trips, DCT scale factors, linearity) cover the whole array.
Tolerances: relative L2 <= 1e-12 (f64), <= 1e-5 (f32)  (BASELINE.json north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from parity_cases import Harness, TOL, cdt  # noqa: E402
from oracle import ndrustfft_oracle as orc  # noqa: E402

OPS = ["ndfft", "ndifft", "ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4"]
TDT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
       np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}


def _backend():
    import ndrustfft_b200 as nb
    be = nb._default_backend()
    assert "sm_100a" in be.lib.version()
    return be


class DevHarness(Harness):
    def __init__(self, be):
        super().__init__(be,
                         mk=lambda a: torch.from_numpy(np.array(a)).cuda(),
                         to_np=lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t),
                         zeros=lambda shape, dt: torch.zeros(tuple(shape), dtype=TDT[np.dtype(dt)], device="cuda"))

    def mk_f(self, a):
        t = torch.from_numpy(np.asfortranarray(np.array(a))).cuda()
        assert t.stride(0) == 1
        return t


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return DevHarness(_backend())


@pytest.fixture(scope="module")
def host():
    return Harness(_backend())  # numpy arrays -> NDFB_MEM_HOST staging path


def test_native_library_is_the_one_running(dev):
    be = dev.be
    before = be.lib.launch_count()
    dev.run("ndfft", 64, (8, 64), 1)
    assert be.lib.launch_count() == before + 1
    assert be.lib.path.endswith("ndrustfft_b200/lib/libndfft_b200.so")


def test_reference_unit_tests_device(dev):
    dev.reference_unit_tests()


def test_reference_unit_tests_host(host):
    host.reference_unit_tests()


LENGTHS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 16, 17, 25, 27, 32, 49, 64, 81, 97, 100, 121, 125, 128, 129, 169,
           243, 256, 264, 343, 360, 384, 500, 512, 513, 625, 1000, 1009, 1024, 1025, 2048, 2187, 4095, 4096]


@pytest.mark.parametrize("n", LENGTHS)
def test_lengths_f64(dev, n):
    for op in OPS:
        if op == "nddct1" and n < 2:
            continue
        dev.run(op, n, (5, n), 1, np.float64, seed=n)
        dev.run(op, n, (n, 6), 0, np.float64, seed=n + 1)


@pytest.mark.parametrize("n", [2, 6, 16, 17, 30, 64, 100, 128, 243, 360, 1000, 1009, 1024, 4096, 8192])
def test_lengths_f32(dev, n):
    for op in OPS:
        dev.run(op, n, (7, n), 1, np.float32, seed=n)
        dev.run(op, n, (n, 9), 0, np.float32, seed=n + 1)


@pytest.mark.parametrize("n", [4099, 5003])
def test_long_bluestein(dev, n):
    dev.run("ndfft", n, (3, n), 1, np.float32, seed=n)
    dev.run("ndifft", n, (n, 3), 0, np.float32, seed=n)
    if n <= 4099:
        dev.run("ndfft", n, (3, n), 1, np.float64, seed=n)


@pytest.mark.parametrize("op", OPS)
@pytest.mark.parametrize("shape,axis", [((40, 33), 0), ((17, 40, 9), 1), ((40, 5, 6), 0), ((4, 5, 40), 2),
                                        ((2, 3, 2, 40, 2), 3), ((40,), 0), ((3, 2, 2, 2, 2, 40), 5), ((3, 2, 40, 2, 2, 2), 2)])
def test_layout_paths(dev, host, op, shape, axis):
    n = shape[axis]
    dev.run(op, n, shape, axis, np.float64, seed=7)
    dev.run(op, n, shape, axis, np.float64, seed=8, order="F")
    host.run(op, n, shape, axis, np.float32, seed=9)


@pytest.mark.parametrize("op", OPS)
def test_norm_none(dev, op):
    dev.run(op, 100, (30, 100), 1, np.float64, norm="none", seed=3)
    dev.run(op, 99, (99, 40), 0, np.float64, norm="none", seed=4)


def test_strided_views_host(host):
    be = host.be
    rng = np.random.default_rng(5)
    base = rng.uniform(-1, 1, (8, 20)) + 1j * rng.uniform(-1, 1, (8, 20))
    x = base[::2, ::-2]
    out_base = np.full((4, 25), 7 + 7j)
    y = out_base[:, 3:23:2]
    be.ndfft(x, y, be.FftHandler(10), 1)
    yo = np.zeros((4, 10), complex)
    orc.ndfft(np.ascontiguousarray(x), yo, orc.FftHandler(10), 1)
    assert orc.rel_l2(y, yo) < 1e-12
    mask = np.ones(25, bool); mask[3:23:2] = False
    assert np.all(out_base[:, mask] == 7 + 7j)


def test_strided_views_device(dev):
    be = dev.be
    rng = np.random.default_rng(6)
    base = rng.uniform(-1, 1, (64, 96))
    t = torch.from_numpy(base).cuda()
    x = t[::2, 1::3]                       # (32, 32) view, strides (192, 3)
    y = torch.zeros((32, 17), dtype=torch.complex128, device="cuda")
    be.ndfft_r2c(x, y, be.R2cFftHandler(32), 1)
    yo = np.zeros((32, 17), complex)
    orc.ndfft_r2c(base[::2, 1::3], yo, orc.R2cFftHandler(32), 1)
    assert orc.rel_l2(y.cpu().numpy(), yo) < 1e-12
    y0 = torch.zeros((17, 32), dtype=torch.complex128, device="cuda")
    be.ndfft_r2c(x, y0, be.R2cFftHandler(32), 0)
    orc_out = np.zeros((17, 32), complex)
    orc.ndfft_r2c(base[::2, 1::3], orc_out, orc.R2cFftHandler(32), 0)
    assert orc.rel_l2(y0.cpu().numpy(), orc_out) < 1e-12


def test_custom_normalization_device(dev):
    be = dev.be
    Norm = type(be.FftHandler(4).norm)

    def f(lane):
        lane *= 0.25
        lane[0] += 1.0

    n = 16
    rng = np.random.default_rng(2)
    sp = rng.uniform(-1, 1, (3, n // 2 + 1)) + 1j * rng.uniform(-1, 1, (3, n // 2 + 1))
    r = torch.zeros((3, n), dtype=torch.float64, device="cuda"); ro = np.zeros((3, n))
    be.ndifft_r2c(torch.from_numpy(sp).cuda(), r, be.R2cFftHandler(n).normalization(Norm.Custom(f)), 1)
    orc.ndifft_r2c(sp, ro, orc.R2cFftHandler(n).normalization(orc.Normalization.custom(f)), 1)
    assert orc.rel_l2(r.cpu().numpy(), ro) < 1e-12
    x = rng.uniform(-1, 1, (3, n)) + 1j * rng.uniform(-1, 1, (3, n))
    y = torch.zeros((3, n), dtype=torch.complex128, device="cuda"); yo = np.zeros((3, n), complex)
    be.ndifft(torch.from_numpy(x).cuda(), y, be.FftHandler(n).normalization(Norm.Custom(f)), 1)
    orc.ndifft(x, yo, orc.FftHandler(n).normalization(orc.Normalization.custom(f)), 1)
    assert orc.rel_l2(y.cpu().numpy(), yo) < 1e-12


def test_errors(dev):
    be = dev.be
    h = be.FftHandler(6)
    z = lambda *s: torch.zeros(s, dtype=torch.complex128, device="cuda")
    with pytest.raises(AssertionError, match="Size mismatch in fft, got 5 expected 6"):
        be.ndfft(z(2, 5), z(2, 5), h, 1)
    with pytest.raises(AssertionError, match="Size mismatch in dct, got 5 expected 6"):
        r = torch.zeros((5, 2), dtype=torch.float64, device="cuda")
        be.nddct3(r, r.clone(), be.DctHandler(6), 0)
    with pytest.raises(AssertionError):
        be.ndfft(z(2, 6), z(3, 6), h, 1)
    with pytest.raises(IndexError):
        be.ndfft(z(2, 6), z(2, 6), h, 2)


# ---------------------------------------------------------------------------------------------------
# BASELINE.json configs at full size
# ---------------------------------------------------------------------------------------------------
def _rand(shape, rd, complex_, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    rt = TDT[np.dtype(rd)]
    if complex_:
        re = torch.rand(shape, generator=g, device="cuda", dtype=rt) * 2 - 1
        im = torch.rand(shape, generator=g, device="cuda", dtype=rt) * 2 - 1
        return torch.complex(re, im)
    return torch.rand(shape, generator=g, device="cuda", dtype=rt) * 2 - 1


def _lane_subset_check(be, op, n, x, y, axis, rd, nsample=48, seed=0):
    """Compare `nsample` random lanes of the full-size result with the oracle."""
    rng = np.random.default_rng(seed)
    xm = torch.movedim(x, axis, -1).reshape(-1, x.shape[axis])
    ym = torch.movedim(y, axis, -1).reshape(-1, y.shape[axis])
    idx = torch.from_numpy(rng.choice(xm.shape[0], size=min(nsample, xm.shape[0]), replace=False)).cuda()
    xs = xm[idx].cpu().numpy()
    ys = ym[idx].cpu().numpy()
    hk = Harness.OPS[op][0]
    yo = np.zeros(ys.shape, np.complex128 if np.iscomplexobj(ys) else np.float64)
    getattr(orc, op)(xs, yo, getattr(orc, hk)(n), 1)
    err = orc.rel_l2(ys, yo)
    assert err <= TOL[np.dtype(rd)], f"{op} n={n} axis={axis}: rel L2 {err:.3e}"
    return err


def _rel(a, b):
    return float((torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item())


def test_config1_128x128_f64(dev):
    # benches/ndrustfft.rs:9-60 shape: n x n f64, ramp data, axis 0; plus axis 1 and random data
    n = 128
    for axis in (0, 1):
        for op in ("ndfft", "ndfft_r2c", "nddct2", "nddct1"):
            dev.run(op, n, (n, n), axis, np.float64, seed=1)
    be = dev.be
    ramp = np.arange(n * n, dtype=np.float64).reshape(n, n)
    y = torch.zeros((n, n), dtype=torch.complex128, device="cuda"); yo = np.zeros((n, n), complex)
    be.ndfft_par(torch.from_numpy(ramp * (1 + 1j)).cuda(), y, be.FftHandler(n), 0)
    orc.ndfft_par(ramp * (1 + 1j), yo, orc.FftHandler(n), 0)
    assert orc.rel_l2(y.cpu().numpy(), yo) < 1e-12


def test_config2_8192x8192_c64(dev):
    be = dev.be
    n = 8192
    x = _rand((n, n), np.float32, True, 0xB200 + 32)
    h = be.FftHandler(n, np.float32)
    y = torch.empty_like(x); z = torch.empty_like(x)
    for axis in (1, 0):
        be.ndfft(x, y, h, axis)
        _lane_subset_check(be, "ndfft", n, x, y, axis, np.float32, seed=axis)
        be.ndifft(y, z, h, axis)
        assert _rel(z, x) < 1e-5                       # round trip over the whole array
        _lane_subset_check(be, "ndifft", n, y, z, axis, np.float32, seed=axis + 2)
    # linearity on the whole array: F(a x + b x') = a F(x) + b F(x')
    x2 = _rand((n, n), np.float32, True, 99)
    be.ndfft(x2, z, h, 0)
    be.ndfft(x, y, h, 0)
    comb = 0.5 * x - 2.0 * x2
    out = torch.empty_like(x)
    be.ndfft(comb, out, h, 0)
    assert _rel(out, 0.5 * y - 2.0 * z) < 1e-5


def test_config3_512cubed_r2c_f64(dev):
    be = dev.be
    n = 512
    x = _rand((n, n, n), np.float64, False, 0xB200 + 48)
    hr = be.R2cFftHandler(n); hc = be.FftHandler(n)
    a = torch.empty((n, n, n // 2 + 1), dtype=torch.complex128, device="cuda")
    b = torch.empty_like(a); c = torch.empty_like(a)
    be.ndfft_r2c(x, a, hr, 2)
    _lane_subset_check(be, "ndfft_r2c", n, x, a, 2, np.float64)
    be.ndfft(a, b, hc, 1)
    _lane_subset_check(be, "ndfft", n, a, b, 1, np.float64)
    be.ndfft(b, c, hc, 0)
    _lane_subset_check(be, "ndfft", n, b, c, 0, np.float64)
    # inverse chain returns the input (examples/rfft2.rs:49-53 pattern)
    be.ndifft(c, b, hc, 0)
    be.ndifft(b, a, hc, 1)
    back = torch.empty_like(x)
    be.ndifft_r2c(a, back, hr, 2)
    assert _rel(back, x) < 1e-12
    # checksum: DC bin of the 3-D spectrum equals the sum of the input
    assert abs(c[0, 0, 0].real.item() - x.sum().item()) <= 1e-9 * n ** 3


def test_config4_4096x4096_dct_f64(dev):
    be = dev.be
    n = 4096
    x = _rand((n, n), np.float64, False, 0xB200 + 64)
    h = be.DctHandler(n)
    y = torch.empty_like(x); z = torch.empty_like(x)
    for axis in (0, 1):
        be.nddct2(x, y, h, axis)
        _lane_subset_check(be, "nddct2", n, x, y, axis, np.float64, nsample=24)
        be.nddct3(y, z, h, axis)
        assert _rel(z, 2 * n * x) < 1e-12
        _lane_subset_check(be, "nddct3", n, y, z, axis, np.float64, nsample=24)
        be.nddct1(x, y, h, axis)
        _lane_subset_check(be, "nddct1", n, x, y, axis, np.float64, nsample=24)
        be.nddct1(y, z, h, axis)
        assert _rel(z, 2 * (n - 1) * x) < 1e-12
        be.nddct4(x, y, h, axis)
        _lane_subset_check(be, "nddct4", n, x, y, axis, np.float64, nsample=24)
        be.nddct4(y, z, h, axis)
        assert _rel(z, 2 * n * x) < 1e-12


def test_config5a_360x1000x384_c128(dev):
    be = dev.be
    shape = (360, 1000, 384)
    x = _rand(shape, np.float64, True, 0xB200 + 80)
    y = torch.empty_like(x); z = torch.empty_like(x)
    for axis in (0, 1, 2):
        h = be.FftHandler(shape[axis])
        be.ndfft(x, y, h, axis)
        _lane_subset_check(be, "ndfft", shape[axis], x, y, axis, np.float64, nsample=32)
        be.ndifft(y, z, h, axis)
        assert _rel(z, x) < 1e-12


def test_config5a_bluestein_axis(dev):
    # the named shape never triggers Bluestein; a 1009-long axis does (SURVEY.md 8d note)
    be = dev.be
    shape = (64, 1009, 48)
    x = _rand(shape, np.float64, True, 5)
    y = torch.empty_like(x); z = torch.empty_like(x)
    h = be.FftHandler(1009)
    assert h.describe()["ops"][0]["family"] == "bluestein"
    be.ndfft(x, y, h, 1)
    _lane_subset_check(be, "ndfft", 1009, x, y, 1, np.float64, nsample=32)
    be.ndifft(y, z, h, 1)
    assert _rel(z, x) < 1e-12


@pytest.mark.parametrize("n,batch", [(1 << 16, 8), (1 << 20, 4), (3 * (1 << 18), 2)])
def test_four_step_medium(dev, n, batch):
    be = dev.be
    x = _rand((batch, n), np.float32, True, n % 1000)
    h = be.FftHandler(n, np.float32)
    y = torch.empty_like(x); z = torch.empty_like(x)
    be.ndfft(x, y, h, 1)
    _lane_subset_check(be, "ndfft", n, x, y, 1, np.float32, nsample=2)
    be.ndifft(y, z, h, 1)
    assert _rel(z, x) < 1e-5
    x64 = _rand((2, n), np.float64, True, 3)
    h64 = be.FftHandler(n, np.float64)
    y64 = torch.empty_like(x64)
    be.ndfft(x64, y64, h64, 1)
    _lane_subset_check(be, "ndfft", n, x64, y64, 1, np.float64, nsample=2)


def test_config5b_2pow24_c64_batch64(dev):
    be = dev.be
    n, batch = 1 << 24, 64
    x = _rand((batch, n), np.float32, True, 0xB200 + 81)
    h = be.FftHandler(n, np.float32)
    assert h.describe()["ops"][0]["family"] == "four-step"
    y = torch.empty_like(x)
    be.ndfft(x, y, h, 1)
    _lane_subset_check(be, "ndfft", n, x, y, 1, np.float32, nsample=2)
    # Parseval on the whole array: sum |X|^2 = n sum |x|^2
    ex = torch.linalg.vector_norm(x).item() ** 2
    ey = torch.linalg.vector_norm(y).item() ** 2
    assert abs(ey / (n * ex) - 1) < 1e-4
    back = torch.empty_like(x)
    be.ndifft(y, back, h, 1)
    assert _rel(back, x) < 1e-5


# ---- staged path at sizes that really overflow one CTA (real kinds, large primes) ----
@pytest.mark.parametrize("op,n,rd", [("ndfft_r2c", 1 << 18, np.float64), ("ndifft_r2c", 1 << 18, np.float32),
                                      ("nddct2", 1 << 16, np.float64), ("nddct3", 1 << 16, np.float64),
                                      ("nddct4", 40000, np.float32), ("nddct1", 32769, np.float64),
                                      ("ndfft", 65537, np.float64), ("ndifft", 1000003, np.float32),
                                      ("ndfft_r2c", 2 * 10007, np.float64),
                                      # odd lengths beyond one CTA: full-length complex core on workspace rows
                                      ("ndfft_r2c", 30011, np.float64), ("ndifft_r2c", 20001, np.float64),
                                      ("nddct2", 20001, np.float64), ("nddct3", 16385, np.float32), ("nddct4", 15001, np.float32)])
def test_staged_long_lanes(dev, op, n, rd):
    dev.run(op, n, (3, n), 1, rd, seed=n % 977)
    if n <= (1 << 16):
        dev.run(op, n, (n, 4), 0, rd, seed=n % 977 + 1)


def test_plan_families_reported(dev):
    be = dev.be
    fam = lambda h: [o["family"] for o in h.describe()["ops"]]
    assert fam(be.FftHandler(8192, np.float32)) == ["direct", "direct"]
    assert fam(be.FftHandler(1 << 24, np.float32))[0] == "four-step"
    assert fam(be.FftHandler(1000003, np.float32))[0] == "staged"
    assert fam(be.R2cFftHandler(1 << 20))[0] == "staged"
    assert fam(be.FftHandler(1009))[0] == "bluestein"


def test_shared_handler_from_threads(dev):
    """One handler used concurrently from several host threads (the reference shares `&handler` across rayon workers,
    src/lib.rs:169-238); every thread gets its own workspaces, results must match the single-threaded ones."""
    import threading
    be = dev.be
    n = 2048
    h = be.FftHandler(n, np.float32)
    xs = [_rand((256, n), np.float32, True, 100 + i) for i in range(4)]
    want = []
    for x in xs:
        y = torch.empty_like(x); be.ndfft(x, y, h, 1); want.append(y)
    outs = [torch.empty_like(x) for x in xs]
    errs = []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(20):
                    be.ndfft(xs[i], outs[i], h, 1)
            s.synchronize()
        except Exception as e:          # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in ts]; [t.join() for t in ts]
    assert not errs
    torch.cuda.synchronize()
    for a, b in zip(outs, want):
        assert torch.equal(a, b)


def test_current_device_is_restored(dev):
    be = dev.be
    before = torch.cuda.current_device()
    dev.run("ndfft", 64, (4, 64), 1)
    assert torch.cuda.current_device() == before


# ---- multi-axis chains (ndfb_exec_chain): same result as the single calls, intermediates on the device ----
def test_chain_examples_device_and_host(dev, host):
    dev.chain_examples()
    host.chain_examples()


@pytest.mark.parametrize("rd", [np.float64, np.float32])
def test_chain_small_device_and_host(dev, host, rd):
    for h in (dev, host):
        h.run_chain([("ndfft", 360, 1), ("ndfft", 100, 0)], (100, 360), rd, seed=1)
        h.run_chain([("ndfft_r2c", 128, 2), ("ndfft", 60, 1), ("ndfft", 36, 0)], (36, 60, 128), rd, seed=2)
        h.run_chain([("ndifft", 36, 0), ("ndifft", 60, 1), ("ndifft_r2c", 128, 2)], (36, 60, 65), rd, seed=3)
        h.run_chain([("nddct1", 129, 1), ("nddct2", 64, 0)], (64, 129), rd, seed=4)
        h.run_chain([("ndfft", 1009, 1), ("ndfft", 17, 0)], (17, 1009), rd, seed=5)
    dev.run_chain([("ndfft", 1024, 1), ("ndfft", 512, 0)], (512, 1024), rd, seed=6, inplace=True)


def test_chain_host_pipelined_matches_single_calls(host, dev):
    # 2048 x 4096 c64 = 64 MiB each way: large enough for the pipelined host path (pieces of the first step behind the
    # upload, pieces of the last step ahead of the download); compare with the two single device calls
    be = host.be
    rng = np.random.default_rng(7)
    x = (rng.uniform(-1, 1, (2048, 4096)) + 1j * rng.uniform(-1, 1, (2048, 4096))).astype(np.complex64)
    y = np.zeros_like(x)
    h0, h1 = be.FftHandler(2048, np.float32), be.FftHandler(4096, np.float32)
    be.fft2(x, y, h0, h1)
    xd = torch.from_numpy(x).cuda(); w = torch.empty_like(xd); yd = torch.empty_like(xd)
    be.ndfft(xd, w, h1, 1); be.ndfft(w, yd, h0, 0)
    assert np.array_equal(y, yd.cpu().numpy())          # same kernels, same order: bit-identical
    want = np.fft.fft2(x.astype(np.complex128))
    assert orc.rel_l2(y, want) <= 1e-5
    # rfft3-style chain with a shape change and a 3-D array (pieces are 2-D copies when axis 0 is transformed last)
    xr = rng.uniform(-1, 1, (128, 256, 512))
    yr = np.zeros((128, 256, 257), np.complex128)
    hr, hb, ha = be.R2cFftHandler(512), be.FftHandler(256), be.FftHandler(128)
    be.ndchain(xr, yr, [("ndfft_r2c", hr, 2), ("ndfft", hb, 1), ("ndfft", ha, 0)])
    assert orc.rel_l2(yr, np.fft.rfftn(xr)) <= 1e-12
    back = np.zeros_like(xr)
    be.ndchain(yr, back, [("ndifft", ha, 0), ("ndifft", hb, 1), ("ndifft_r2c", hr, 2)])
    assert orc.rel_l2(back, xr) <= 1e-12


def test_chain_config3_device(dev):
    # c3 as one call: 512^3 f64 -> 512 x 512 x 257 c128, intermediates live in the output array (two in-place passes)
    be = dev.be
    n = 512
    x = _rand((n, n, n), np.float64, False, 0xB200 + 49)
    hr = be.R2cFftHandler(n); hc = be.FftHandler(n)
    c = torch.empty((n, n, n // 2 + 1), dtype=torch.complex128, device="cuda")
    be.ndchain(x, c, [("ndfft_r2c", hr, 2), ("ndfft", hc, 1), ("ndfft", hc, 0)])
    a = torch.empty_like(c); b = torch.empty_like(c)
    be.ndfft_r2c(x, a, hr, 2); be.ndfft(a, b, hc, 1); be.ndfft(b, a, hc, 0)
    assert torch.equal(a, c)
    back = torch.empty_like(x)
    be.ndchain(c, back, [("ndifft", hc, 0), ("ndifft", hc, 1), ("ndifft_r2c", hr, 2)])
    assert _rel(back, x) < 1e-12


@pytest.mark.parametrize("op,n,shape,axis", [("ndfft", 8192, (64, 8192), 1), ("ndifft", 8192, (8192, 64), 0), ("ndfft", 1 << 20, (2, 1 << 20), 1),
                                              ("nddct1", 4096, (4096, 32), 0), ("nddct3", 4096, (16, 4096), 1), ("ndfft", 1009, (1009, 40), 0)])
def test_single_call_in_place_device(dev, op, n, shape, axis):
    be = dev.be
    icx = Harness.OPS[op][1]
    rd = np.float32 if n >= 8192 else np.float64
    x = _rand(shape, rd, icx, 11)
    h = getattr(be, Harness.OPS[op][0])(n, rd)
    want = torch.empty_like(x)
    getattr(be, op)(x, want, h, axis)
    buf = x.clone()
    getattr(be, op)(buf, buf, h, axis)
    assert torch.equal(buf, want)


def test_fused_two_pass_columns_device(dev):
    # opt-in (NDFB_FS2=1): both column passes of c2 axis 0 in one persistent launch (fs2_kernel); same bits as the default
    # two-launch path
    import os
    be = dev.be
    x = _rand((8192, 1024), np.float32, True, 21)
    h = be.FftHandler(8192, np.float32)
    y = torch.empty_like(x); y2 = torch.empty_like(x)
    l0 = be.lib.launch_count()
    be.ndfft(x, y2, h, 0)
    assert be.lib.launch_count() - l0 == 2
    os.environ["NDFB_FS2"] = "1"
    os.environ["NDFB_FS2_F64"] = "1"
    try:
        l0 = be.lib.launch_count()
        be.ndfft(x, y, h, 0)
        assert be.lib.launch_count() - l0 == 1
        assert torch.equal(y, y2)
        _lane_subset_check(be, "ndfft", 8192, x, y, 0, np.float32, nsample=16)
        # inverse, in place, and a column count that is not a power of two (groups of 192 = 3 x 64 columns)
        xi = _rand((8192, 960), np.float32, True, 22)
        buf = xi.clone()
        be.ndfft(buf, buf, h, 0); be.ndifft(buf, buf, h, 0)
        assert _rel(buf, xi) < 1e-5
        # double precision on request
        xd = _rand((8192, 512), np.float64, True, 23)
        hd = be.FftHandler(8192, np.float64)
        yd = torch.empty_like(xd)
        l0 = be.lib.launch_count()
        be.ndfft(xd, yd, hd, 0)
        assert be.lib.launch_count() - l0 == 1
        _lane_subset_check(be, "ndfft", 8192, xd, yd, 0, np.float64, nsample=16)
    finally:
        del os.environ["NDFB_FS2_F64"]
        del os.environ["NDFB_FS2"]


def test_workspace_reuse_across_streams(dev):
    # two-pass transforms from one host thread on two streams share the thread's workspace: the library orders them
    be = dev.be
    h = be.FftHandler(1 << 16, np.float32)
    x1 = _rand((64, 1 << 16), np.float32, True, 31); x2 = _rand((64, 1 << 16), np.float32, True, 32)
    w1 = torch.empty_like(x1); w2 = torch.empty_like(x2)
    be.ndfft(x1, w1, h, 1); be.ndfft(x2, w2, h, 1)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    y1 = torch.empty_like(x1); y2 = torch.empty_like(x2)
    for _ in range(5):
        with torch.cuda.stream(s1):
            be.ndfft(x1, y1, h, 1)
        with torch.cuda.stream(s2):
            be.ndfft(x2, y2, h, 1)
    torch.cuda.synchronize()
    assert torch.equal(y1, w1) and torch.equal(y2, w2)


def test_three_pass_split_matches_two_pass(dev, capfd):
    # 2^24-point rows: the default is three passes over 256-byte rows, 256 x (256 x 256); forced 4096 x 4096 (two passes over 32-byte
    # rows) and 128 x (...): the nested level's last pass tiles the lanes along the output-contiguous dim (transposing pass).
    # Same transform, different factorisations: equal within f32 rounding.
    import os
    be = dev.be
    n = 1 << 24
    x = _rand((2, n), np.float32, True, 41)
    h = be.FftHandler(n, np.float32)
    y2 = torch.empty_like(x); y3 = torch.empty_like(x); yd = torch.empty_like(x)
    os.environ["NDFB_TRACE"] = "1"
    try:
        be.ndfft(x, yd, h, 1)
        assert "= 256 x 65536" in capfd.readouterr().err
        os.environ["NDFB_FS_N1"] = "4096"; os.environ["NDFB_FS_TWO_PASS"] = "1"
        be.ndfft(x, y2, h, 1)
        assert "= 4096 x 4096 (contiguous lanes)" in capfd.readouterr().err
        del os.environ["NDFB_FS_TWO_PASS"]
        os.environ["NDFB_FS_N1"] = "128"
        be.ndfft(x, y3, h, 1)
        assert _rel(y3, y2) < 2e-6 and _rel(yd, y2) < 2e-6
        be.ndifft(y3, y3, h, 1)          # in place through the three-pass path
        assert _rel(y3, x) < 2e-6
    finally:
        os.environ.pop("NDFB_FS_N1", None); os.environ.pop("NDFB_TRACE", None); os.environ.pop("NDFB_FS_TWO_PASS", None)
    _lane_subset_check(be, "ndfft", n, x, yd, 1, np.float32, nsample=1)


# ---- round-2 additions: the holes VERDICT r1 listed ----
def test_negative_and_reversed_strides_device(dev):
    """Device views with negative strides (ndarray's slice(s![..;-1]) / invert_axis): base pointer = element [0,...], the
    kernels walk backwards through memory."""
    be = dev.be
    rng = np.random.default_rng(16)
    base = rng.uniform(-1, 1, (48, 64)) + 1j * rng.uniform(-1, 1, (48, 64))
    t = torch.from_numpy(base).cuda()
    x = t.flip(0)            # torch has no negative strides: build the reversed view through the C ABI by hand below
    import ctypes
    from ndrustfft_b200 import _lib
    for axis, n in ((1, 64), (0, 48)):
        h = be.FftHandler(n)
        y = torch.zeros((48, 64), dtype=torch.complex128, device="cuda")
        yv = torch.zeros((48, 64), dtype=torch.complex128, device="cuda")
        SZ, PD = ctypes.c_size_t * 2, ctypes.c_ssize_t * 2
        # input: both axes reversed (element [0,0] = last element of the buffer); output: rows reversed
        in_ptr = t.data_ptr() + (48 * 64 - 1) * 16
        out_ptr = yv.data_ptr() + (47 * 64) * 16
        rc = be.lib.dll.ndfb_exec(h._plan, _lib.OP_FFT, _lib.NORM_DEFAULT, ctypes.c_void_p(in_ptr), ctypes.c_void_p(out_ptr), 2,
                                  SZ(48, 64), PD(-64, -1), SZ(48, 64), PD(-64, 1), axis, _lib.MEM_DEVICE,
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        be.lib.check(rc)
        want = np.zeros((48, 64), complex)
        orc.ndfft(np.ascontiguousarray(base[::-1, ::-1]), want, orc.FftHandler(n), axis)
        got = yv.cpu().numpy()[::-1, :]          # logical output [i, j] lives at physical row 47 - i
        assert orc.rel_l2(got, want) < 1e-12, axis
    # real kinds with a reversed transformed axis, strided columns
    xr = rng.uniform(-1, 1, (128, 40))
    tr = torch.from_numpy(xr).cuda()
    yr = torch.zeros((128, 40), dtype=torch.float64, device="cuda")
    h = be.DctHandler(128)
    SZ, PD = ctypes.c_size_t * 2, ctypes.c_ssize_t * 2
    rc = be.lib.dll.ndfb_exec(h._plan, _lib.OP_DCT2, _lib.NORM_DEFAULT, ctypes.c_void_p(tr.data_ptr() + 127 * 40 * 8), ctypes.c_void_p(yr.data_ptr()), 2,
                              SZ(128, 40), PD(-40, 1), SZ(128, 40), PD(40, 1), 0, _lib.MEM_DEVICE, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    be.lib.check(rc)
    want = np.zeros((128, 40))
    orc.nddct2(np.ascontiguousarray(xr[::-1]), want, orc.DctHandler(128), 0)
    assert orc.rel_l2(yr.cpu().numpy(), want) < 1e-12
    del x


def test_scatter_out_single_gpu(dev):
    """ndfb_exec_scatter_out on ONE GPU: the block pointers address four regions of one buffer, so the driver's 1-GPU test
    run covers the store path the multi-GPU slab exchange uses (peer-mapped buffers there)."""
    be = dev.be
    rng = np.random.default_rng(17)
    s0, n1, mp, P = 6, 512, 264, 4
    s1 = n1 // P
    x = rng.uniform(-1, 1, (s0, n1, mp)) + 1j * rng.uniform(-1, 1, (s0, n1, mp))
    xd = torch.from_numpy(x).cuda()
    recv = torch.zeros((P, s0, s1, mp), dtype=torch.complex128, device="cuda")
    ptrs = [recv.data_ptr() + p * s0 * s1 * mp * 16 for p in range(P)]
    be.ndfft_scatter_out(xd, be.FftHandler(n1), 1, out_shape=(s0, n1, mp), out_strides=(s1 * mp, mp, 1), out_block=s1, block_ptrs=ptrs)
    want = np.fft.fft(x, axis=1)
    got = recv.cpu().numpy()
    for p in range(P):
        assert orc.rel_l2(got[p], want[:, p * s1:(p + 1) * s1, :]) < 1e-12
    # i2-chunked form (the overlapped pipeline scatters sub-ranges of the last dim)
    recv.zero_()
    for lo, hi in ((0, 128), (128, 264)):
        be.ndfft_scatter_out(xd[:, :, lo:hi], be.FftHandler(n1), 1, out_shape=(s0, n1, hi - lo), out_strides=(s1 * mp, mp, 1), out_block=s1,
                             block_ptrs=[q + lo * 16 for q in ptrs])
    got = recv.cpu().numpy()
    for p in range(P):
        assert orc.rel_l2(got[p], want[:, p * s1:(p + 1) * s1, :]) < 1e-12


def test_pageable_host_arrays_pipelined(host, dev):
    """Pageable numpy arrays (what ndarray hands the shim) through the pinned staging ring: bit-identical to the device path,
    for a contiguous-axis call (1-D pieces), a strided-axis call (2-D pieces) and an r2c (shape change)."""
    be = host.be
    rng = np.random.default_rng(18)
    x = (rng.uniform(-1, 1, (4096, 4096)) + 1j * rng.uniform(-1, 1, (4096, 4096))).astype(np.complex64)     # 128 MiB
    h = be.FftHandler(4096, np.float32)
    xd = torch.from_numpy(x).cuda()
    for axis in (1, 0):
        y = np.zeros_like(x)
        be.ndfft(x, y, h, axis)
        yd = torch.empty_like(xd)
        be.ndfft(xd, yd, h, axis)
        assert np.array_equal(y, yd.cpu().numpy()), axis
    xr = rng.uniform(-1, 1, (3000, 4096))
    yr = np.zeros((3000, 2049), np.complex128)
    be.ndfft_r2c(xr, yr, be.R2cFftHandler(4096), 1)
    assert orc.rel_l2(yr, np.fft.rfft(xr, axis=1)) < 1e-12
    # gappy views of large arrays: packed on the host, pipelined, unpacked; the gaps stay intact
    big = np.full((2048, 2 * 4096), 1.5 - 0.5j, np.complex64)
    out = big[:, ::2]
    be.ndfft(x[:2048], out, h, 1)
    yd = torch.empty((2048, 4096), dtype=torch.complex64, device="cuda")
    be.ndfft(xd[:2048], yd, h, 1)
    assert np.array_equal(np.ascontiguousarray(out), yd.cpu().numpy())
    assert np.all(big[:, 1::2] == np.complex64(1.5 - 0.5j))


def test_config1_in_cuda_graph(dev):
    """c1 is launch-latency bound: the calls must be capturable into a CUDA graph (plan tables uploaded beforehand)."""
    be = dev.be
    rng = np.random.default_rng(19)
    n = 128
    x = rng.uniform(-1, 1, (n, n))
    xd = torch.from_numpy(x).cuda()
    hr, hc, hd = be.R2cFftHandler(n), be.FftHandler(n), be.DctHandler(n)
    a = torch.zeros((n, n // 2 + 1), dtype=torch.complex128, device="cuda")
    b = torch.zeros_like(a)
    c = torch.zeros((n, n), dtype=torch.float64, device="cuda")

    def calls():
        be.ndfft_r2c(xd, a, hr, 1); be.ndfft(a, b, hc, 0); be.nddct2(xd, c, hd, 0)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        calls()                                   # warm-up outside the capture: tables, function attributes
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            calls()
    torch.cuda.current_stream().wait_stream(side)
    b.zero_(); c.zero_()
    g.replay()
    torch.cuda.synchronize()
    want = np.fft.fft(np.fft.rfft(x, axis=1), axis=0)
    assert orc.rel_l2(b.cpu().numpy(), want) < 1e-12
    wc = np.zeros((n, n)); orc.nddct2(x, wc, orc.DctHandler(n), 0)
    assert orc.rel_l2(c.cpu().numpy(), wc) < 1e-12


# ---- run-time compiled schedules (csrc/jit.h): smooth lengths without an ahead-of-time instance ----
@pytest.mark.parametrize("n", [96, 192, 720, 768, 1200, 1536, 3072, 6561, 1001])
def test_jit_lengths_c2c(dev, n, capfd):
    import os
    os.environ["NDFB_TRACE"] = "1"
    try:
        dev.run("ndfft", n, (6, n), 1, np.float64, seed=n)
        dev.run("ndifft", n, (n, 12), 0, np.float32, seed=n + 1)
        dev.run("ndfft", n, (3, n, 5), 1, np.float64, seed=n + 2, norm="none")
    finally:
        del os.environ["NDFB_TRACE"]
    err = capfd.readouterr().err
    if "run-time schedule compilation unavailable" in err:
        pytest.skip("libnvrtc not usable on this box")
    # sfft_kernel instances compiled at run time, not the general tile kernel (6561-point columns do not fit one CTA four
    # lanes wide: they run as 81 x 81 two-pass transforms on instantiated schedules)
    assert err.count(" jit ") >= (1 if n == 6561 else 3), err
    assert "tile_kernel" not in err


@pytest.mark.parametrize("op", ["ndfft_r2c", "ndifft_r2c", "nddct1", "nddct2", "nddct3", "nddct4"])
@pytest.mark.parametrize("n", [96, 720, 1536])
def test_jit_lengths_real_kinds(dev, op, n, capfd):
    import os
    nn = n + 1 if op == "nddct1" else n
    os.environ["NDFB_TRACE"] = "1"
    try:
        dev.run(op, nn, (5, nn), 1, np.float64, seed=n)
        dev.run(op, nn, (nn, 9), 0, np.float32, seed=n + 1)
    finally:
        del os.environ["NDFB_TRACE"]
    err = capfd.readouterr().err
    if "run-time schedule compilation unavailable" in err:
        pytest.skip("libnvrtc not usable on this box")
    assert err.count("[ndfb] rsfft") == 2 and err.count(" jit") >= 2, err


# ---- round-2 kernels and entry points on the GPU ----
def test_rows_bulk_async_kernel_gpu(dev, capfd):
    """The persistent TMA + mbarrier row kernel (opt-in) gives the bits of the register-resident kernel."""
    import os
    be = dev.be
    rng = np.random.default_rng(21)
    for n, rd, lanes in ((8192, np.float32, 700), (2048, np.float64, 1300), (1024, np.float32, 2051)):
        x = (rng.uniform(-1, 1, (lanes, n)) + 1j * rng.uniform(-1, 1, (lanes, n))).astype(cdt(rd))
        xd = torch.from_numpy(x).cuda()
        y0 = torch.empty_like(xd); y1 = torch.empty_like(xd)
        h = be.FftHandler(n, rd)
        be.ndfft(xd, y0, h, 1)
        os.environ["NDFB_ROWS_BULK"] = "2"; os.environ["NDFB_TRACE"] = "1"
        try:
            be.ndfft(xd, y1, h, 1)
            be.ndifft(y1, xd, h, 1)
        finally:
            del os.environ["NDFB_ROWS_BULK"]; del os.environ["NDFB_TRACE"]
        assert capfd.readouterr().err.count("rows bulk-async persistent") == 2
        assert torch.equal(y0, y1)
        assert orc.rel_l2(xd.cpu().numpy(), x) <= TOL[np.dtype(rd)]


def test_pipelined_column_kernel_gpu(dev, capfd):
    """The software-pipelined persistent column kernel (cp.async staging + split exchange) gives the bits of the register-resident
    kernel: lane-adjacent input (strided axis), both precisions, the tile shapes of c5b / c2 / long f64 columns; then a two-pass
    split whose both passes take it (pass 2 = contiguous workspace rows in, lane-interleaved out)."""
    import os
    be = dev.be
    rng = np.random.default_rng(22)
    for n, rd, lanes in ((4096, np.float32, 1200), (8192, np.float32, 600), (2048, np.float64, 1184), (4096, np.float64, 596), (512, np.float32, 64)):
        x = (rng.uniform(-1, 1, (n, lanes)) + 1j * rng.uniform(-1, 1, (n, lanes))).astype(cdt(rd))
        xd = torch.from_numpy(x).cuda()
        y0 = torch.empty_like(xd); y1 = torch.empty_like(xd)
        h = be.FftHandler(n, rd)
        os.environ["NDFB_PIPE"] = "0"; os.environ["NDFB_STRIDED_FOURSTEP"] = "0"
        try:
            be.ndfft(xd, y0, h, 0)
            os.environ["NDFB_PIPE"] = "2"; os.environ["NDFB_TRACE"] = "1"
            be.ndfft(xd, y1, h, 0)
            be.ndifft(y1, xd, h, 0)
        finally:
            for k in ("NDFB_PIPE", "NDFB_TRACE", "NDFB_STRIDED_FOURSTEP"):
                os.environ.pop(k, None)
        assert capfd.readouterr().err.count("cols pipelined") == 2
        assert torch.equal(y0, y1)
        assert orc.rel_l2(xd.cpu().numpy(), x) <= TOL[np.dtype(rd)]
    # one-lane tiles: contiguous rows in and out (opt-in)
    for n, rd, lanes in ((8192, np.float32, 601), (4096, np.float64, 333)):
        x = (rng.uniform(-1, 1, (lanes, n)) + 1j * rng.uniform(-1, 1, (lanes, n))).astype(cdt(rd))
        xd = torch.from_numpy(x).cuda()
        y0 = torch.empty_like(xd); y1 = torch.empty_like(xd)
        h = be.FftHandler(n, rd)
        os.environ["NDFB_PIPE"] = "0"
        try:
            be.ndfft(xd, y0, h, 1)
            os.environ["NDFB_PIPE"] = "2"; os.environ["NDFB_TRACE"] = "1"
            be.ndfft(xd, y1, h, 1)
        finally:
            for k in ("NDFB_PIPE", "NDFB_TRACE"):
                os.environ.pop(k, None)
        assert capfd.readouterr().err.count("rows pipelined") == 1
        assert torch.equal(y0, y1)
    n = 1 << 24
    x = _rand((6, n), np.float32, True, 77)
    h = be.FftHandler(n, np.float32)
    y0 = torch.empty_like(x); y1 = torch.empty_like(x)
    os.environ.update({"NDFB_PIPE": "0", "NDFB_FS_N1": "4096", "NDFB_FS_TWO_PASS": "1", "NDFB_NO_FS_TRANSPOSE": "1"})     # the two-pass split: 4 x 4096 tiles in both passes
    try:
        be.ndfft(x, y0, h, 1)
        os.environ["NDFB_PIPE"] = "2"; os.environ["NDFB_TRACE"] = "1"
        be.ndfft(x, y1, h, 1)
    finally:
        for k in ("NDFB_PIPE", "NDFB_TRACE", "NDFB_FS_N1", "NDFB_FS_TWO_PASS", "NDFB_NO_FS_TRANSPOSE"):
            os.environ.pop(k, None)
    err = capfd.readouterr().err
    assert err.count("in=lane-adjacent") == 1 and err.count("in=rows") == 1, err
    # pass 1 forms the four-step twiddle as (one lookup per butterfly) x (per-lane factor table): same value, other rounding
    assert float(torch.linalg.vector_norm((y0 - y1).flatten()) / torch.linalg.vector_norm(y0.flatten())) < 2e-7
    _lane_subset_check(be, "ndfft", n, x, y1, 1, np.float32, nsample=2)


@pytest.mark.parametrize("op", ["nddct3", "nddct4"])
@pytest.mark.parametrize("n,rd", [(4096, np.float64), (2048, np.float64), (4096, np.float32), (1536, np.float64)])
def test_mirror_paired_output_pass_gpu(dev, op, n, rd, capfd):
    """DCT-III / DCT-IV rows: aligned contiguous rows take the mirror-paired last pass (pairs of reals stored from registers), rows of
    a device view that start on odd elements keep the staged copy-out; both against the oracle and against each other."""
    import os
    be = dev.be
    rng = np.random.default_rng(n)
    lanes = 37
    x = rng.uniform(-1, 1, (lanes, n)).astype(rd)
    ho = orc.DctHandler(n)
    want = np.zeros((lanes, n)); getattr(orc, op)(x.astype(np.float64), want, ho, 1)
    h = be.DctHandler(n, rd)
    xd = torch.from_numpy(x).cuda()
    os.environ["NDFB_TRACE"] = "1"
    os.environ["NDFB_MIRROR_OUT"] = "1"      # (no longer needed: every DCT-III / DCT-IV row call takes it)
    try:
        y = torch.empty_like(xd)
        getattr(be, op)(xd, y, h, 1)
        err = capfd.readouterr().err
        assert "mirror-paired output" in err
        big_in = torch.zeros((lanes, n + 1), dtype=xd.dtype, device="cuda"); big_in[:, :n] = xd
        big_out = torch.full((lanes, n + 1), 7.0, dtype=xd.dtype, device="cuda")
        getattr(be, op)(big_in[:, :n], big_out[:, :n], h, 1)
        assert "mirror-paired output" not in capfd.readouterr().err
    finally:
        del os.environ["NDFB_TRACE"]; del os.environ["NDFB_MIRROR_OUT"]
    assert orc.rel_l2(y.cpu().numpy(), want) <= TOL[np.dtype(rd)]
    assert orc.rel_l2(big_out[:, :n].cpu().numpy(), want) <= TOL[np.dtype(rd)]
    assert bool((big_out[:, n] == 7.0).all())
    assert orc.rel_l2(y.cpu().numpy(), big_out[:, :n].cpu().numpy()) <= TOL[np.dtype(rd)] * 0.1


def test_device_memory_and_stream_helpers(dev):
    """ndfb_device_alloc / ndfb_memcpy / ndfb_stream_*: what the Rust shim's DeviceArray and Stream are built on; pageable
    uploads and downloads above 4 MiB go through the pinned ring."""
    import ctypes
    from ndrustfft_b200 import _lib
    be = dev.be
    dll = be.lib.dll
    rng = np.random.default_rng(22)
    n, lanes = 1024, 1500                                   # 24 MiB c128: ring path
    x = rng.uniform(-1, 1, (lanes, n)) + 1j * rng.uniform(-1, 1, (lanes, n))
    y = np.zeros_like(x)
    st = ctypes.c_void_p(); din = ctypes.c_void_p(); dout = ctypes.c_void_p()
    be.lib.check(dll.ndfb_stream_create(ctypes.byref(st), 0))
    be.lib.check(dll.ndfb_device_alloc(ctypes.byref(din), x.nbytes, 0))
    be.lib.check(dll.ndfb_device_alloc(ctypes.byref(dout), x.nbytes, 0))
    be.lib.check(dll.ndfb_memcpy(din, ctypes.c_void_p(x.ctypes.data), x.nbytes, 0, 0, st))
    h = be.FftHandler(n)
    SZ, PD = ctypes.c_size_t * 2, ctypes.c_ssize_t * 2
    be.lib.check(dll.ndfb_exec(h._plan, _lib.OP_FFT, _lib.NORM_DEFAULT, din, dout, 2, SZ(lanes, n), PD(n, 1), SZ(lanes, n), PD(n, 1), 1, _lib.MEM_DEVICE, st))
    be.lib.check(dll.ndfb_memcpy(ctypes.c_void_p(y.ctypes.data), dout, x.nbytes, 1, 0, st))
    be.lib.check(dll.ndfb_stream_sync(st))
    assert orc.rel_l2(y, np.fft.fft(x, axis=1)) < 1e-12
    dll.ndfb_device_free(din); dll.ndfb_device_free(dout); dll.ndfb_stream_destroy(st)
