"""Host-array plumbing and many-dimensional views (CPU, SIMT-emulation build of the same sources).

Covers the round-1 advisor findings: the multi-pass paths with more batch dims than their kernels index, host views
with gaps (the library may neither read nor write bytes between the logical elements of a view — the reference only
touches lane elements, src/lib.rs:119-163), scattered output blocks under dim peeling, and the Python mirror's
argument checks."""
import ctypes
import os

import numpy as np
import pytest

from emu_backend import emu_backend
from oracle import ndrustfft_oracle as orc
from ndrustfft_b200 import _lib


@pytest.fixture(scope="module")
def be():
    return emu_backend()


def _cx(rng, shape, dt=np.complex128):
    return (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(dt)


def _env(**kw):
    class E:
        def __enter__(self):
            self.old = {k: os.environ.get(k) for k in kw}
            os.environ.update({k: str(v) for k, v in kw.items()})

        def __exit__(self, *a):
            for k, v in self.old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return E()


def test_four_step_on_6d_sliced_view(be, capfd):
    """ndfft along axis 0 of a 6-D sliced view whose length takes the two-pass path (8192-point c64 on the
    GPU; here 64 points with the 'chip' shrunk by NDFB_FS_CAP) used to fail with 'more than 3 batch dims'."""
    rng = np.random.default_rng(1)
    n = 64
    base = _cx(rng, (n, 3, 4, 3, 4, 5), np.complex64)
    x = base[:, ::2, 1:3, ::2, 1:4:2, 1:4]          # 5 non-mergeable batch dims
    assert x.shape == (n, 2, 2, 2, 2, 3)
    y = np.zeros(x.shape, np.complex64)
    with _env(NDFB_FORCE_FOUR_STEP=1, NDFB_FS_CAP=16, NDFB_TRACE=1):
        be.ndfft(x, y, be.FftHandler(n, np.float32), 0)
        # the same through the DEVICE entry on the strided view itself (host path packs views with gaps first)
        y2 = np.zeros(x.shape, np.complex64)
        h = be.FftHandler(n, np.float32)
        SZ, PD = ctypes.c_size_t * 6, ctypes.c_ssize_t * 6
        rc = be.lib.dll.ndfb_exec(h._plan, _lib.OP_FFT, _lib.NORM_DEFAULT, ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(y2.ctypes.data), 6,
                                  SZ(*x.shape), PD(*[s // 8 for s in x.strides]), SZ(*y2.shape), PD(*[s // 8 for s in y2.strides]), 0,
                                  _lib.MEM_DEVICE, None)
        be.lib.check(rc)
    err = capfd.readouterr().err
    assert "four-step" in err
    want = np.fft.fft(x.astype(np.complex128), axis=0)
    assert orc.rel_l2(y, want) < 1e-5
    assert orc.rel_l2(y2, want) < 1e-5


def test_staged_path_on_7d_view(be):
    rng = np.random.default_rng(2)
    n = 45
    base = rng.uniform(-1, 1, (2, 3, 2, n, 3, 2, 4))
    x = base[:, ::2, :, :, ::2, :, ::3]              # 6 batch dims, non mergeable
    y = np.zeros(x.shape)
    yo = np.zeros(x.shape)
    with _env(NDFB_FORCE_STAGED=1):
        h = be.DctHandler(n)
        SZ, PD = ctypes.c_size_t * 7, ctypes.c_ssize_t * 7
        rc = be.lib.dll.ndfb_exec(h._plan, _lib.OP_DCT2, _lib.NORM_DEFAULT, ctypes.c_void_p(x.ctypes.data), ctypes.c_void_p(y.ctypes.data), 7,
                                  SZ(*x.shape), PD(*[s // 8 for s in x.strides]), SZ(*y.shape), PD(*[s // 8 for s in y.strides]), 3,
                                  _lib.MEM_DEVICE, None)
        be.lib.check(rc)
    orc.nddct2(np.ascontiguousarray(x), yo, orc.DctHandler(n), 3)
    assert orc.rel_l2(y, yo) < 1e-12


def test_host_output_gaps_are_never_written(be):
    """A non-dense host output (e.g. one half of ndarray's multi_slice_mut) shares its byte span with sibling views:
    only the logical elements may change."""
    rng = np.random.default_rng(3)
    n = 32
    x = _cx(rng, (6, n))
    big = np.full((12, 2 * n + 3), 7.5 - 2.5j)
    out = big[1::2, 2:2 + 2 * n:2]
    assert out.shape == (6, n)
    be.ndfft(x, out, be.FftHandler(n), 1)
    want = np.fft.fft(x, axis=1)
    assert orc.rel_l2(out, want) < 1e-12
    mask = np.ones(big.shape, bool)
    mask[1::2, 2:2 + 2 * n:2] = False
    assert np.all(big[mask] == 7.5 - 2.5j), "bytes between the view's elements were overwritten"
    # r2c into a gappy output, from a gappy + reversed input
    xr_base = rng.uniform(-1, 1, (n * 3, 5))
    xr = xr_base[::-3, 1:4]                            # negative stride along the axis, gaps on both dims
    bigc = np.full((n // 2 + 1, 9), -1.0 + 1.0j)
    oc = bigc[:, ::3]
    be.ndfft_r2c(xr, oc, be.R2cFftHandler(n), 0)
    assert orc.rel_l2(oc, np.fft.rfft(xr, axis=0)) < 1e-12
    m2 = np.ones(bigc.shape, bool); m2[:, ::3] = False
    assert np.all(bigc[m2] == -1.0 + 1.0j)


def test_host_chain_gappy_views(be):
    rng = np.random.default_rng(4)
    base = rng.uniform(-1, 1, (10, 24))
    x = base[::2, ::3]                                # (5, 8) real
    big = np.full((5, 2 * 5), 3.0 + 4.0j)
    out = big[:, ::2]                                 # (5, 5) complex, gaps
    be.rfft2(x, out, be.FftHandler(5), be.R2cFftHandler(8))
    want = np.fft.fft(np.fft.rfft(x, axis=1), axis=0)
    assert orc.rel_l2(out, want) < 1e-12
    assert np.all(big[:, 1::2] == 3.0 + 4.0j)


def test_scatter_out_with_peeled_batch_dims(be):
    """with more than 4 batch dims the peel loop must move the scattered block pointers along too."""
    rng = np.random.default_rng(5)
    n1, P = 64, 2
    s1 = n1 // P
    shape = (2, 3, n1, 2, 3, 2, 2)                    # 6 batch dims
    base = _cx(rng, (2, 3, n1, 2, 3, 2, 4))
    x = base[..., ::2]                                # last dim strided so nothing merges with it
    x = x[:, ::-1]                                    # and a reversed dim
    assert x.shape == shape
    oshape = (2, 3, s1, 2, 3, 2, 2)
    bufs = [np.zeros(oshape, complex) for _ in range(P)]
    ostr = [s // 16 for s in bufs[0].strides]
    be.ndfft_scatter_out(x, be.FftHandler(n1), 2, out_shape=shape, out_strides=tuple(ostr), out_block=s1,
                         block_ptrs=[b.ctypes.data for b in bufs])
    want = np.fft.fft(x, axis=2)
    for p in range(P):
        assert orc.rel_l2(bufs[p], want[:, :, p * s1:(p + 1) * s1]) < 1e-12


def test_python_mirror_rejects_bad_arrays(be):
    import torch
    n = 8
    h = be.FftHandler(n)
    x = np.zeros((2, n), complex); y = np.zeros((2, n), complex)
    y.setflags(write=False)
    with pytest.raises(ValueError, match="read-only"):
        be.ndfft(x, y, h, 1)
    tx = torch.zeros((2, n), dtype=torch.complex128)
    ty = torch.zeros((2, n), dtype=torch.complex128)
    with pytest.raises(ValueError, match="conj"):
        be.ndfft(tx.conj(), ty, h, 1)
    # shape-changing ops cannot run in place
    hr = be.R2cFftHandler(n)
    buf = np.zeros(2 * (n // 2 + 1) * 2)
    xr = buf[:2 * n].reshape(2, n)
    yc = buf.view(complex)[:2 * (n // 2 + 1)].reshape(2, n // 2 + 1)
    with pytest.raises(ValueError, match="overlap"):
        be.ndfft_r2c(xr, yc, hr, 1)


def test_blocked_exchange_layout_is_a_stride_description(be):
    """dist.SlabR2cFft3d's blocked receive layout [i0][i2 block][j1][lane]: the scatter store and the last pass only see
    different strides.  Checked here on one process: 'peers' are regions of one buffer."""
    rng = np.random.default_rng(6)
    s0, n1, lb, nb, P = 2, 64, 8, 3, 4
    mp, s1 = lb * nb, n1 // P
    n0 = s0 * P
    a = _cx(rng, (s0, n1, mp))
    # every 'rank' sends the same local array; rank r's planes land in chunk r of each destination buffer
    bufs = [np.zeros(P * s0 * s1 * mp, complex) for _ in range(P)]
    h1, h0 = be.FftHandler(n1), be.FftHandler(n0)
    chunk = s0 * s1 * mp * 16
    for r in range(P):
        be.ndfft_scatter_out(a.reshape(s0, n1, nb, lb), h1, 1, out_shape=(s0, n1, nb, lb), out_strides=(s1 * mp, lb, s1 * lb, 1),
                             out_block=s1, block_ptrs=[b.ctypes.data + r * chunk for b in bufs])
    fa = np.fft.fft(a, axis=1)
    for d in range(P):
        recv = bufs[d].reshape(n0, nb, s1, lb).transpose(0, 2, 1, 3)           # logical (i0, j1, i2 block, lane)
        want = np.concatenate([fa[:, d * s1:(d + 1) * s1, :]] * P, axis=0).reshape(n0, s1, nb, lb)
        assert orc.rel_l2(recv, want) < 1e-12
        out = np.zeros((n0, s1, mp), complex)
        be.ndfft(recv, out.reshape(n0, s1, nb, lb), h0, 0)
        assert orc.rel_l2(out, np.fft.fft(want.reshape(n0, s1, mp), axis=0)) < 1e-12


def test_bulk_async_scatter_blocks(be, capfd):
    """512-point c128 columns, tile 8 lanes wide, blocked layout: each (tile, destination) block is contiguous, so the kernel
    stages the tile and sends every block with one bulk copy (cp.async.bulk on the GPU; a byte loop under the emulator)."""
    rng = np.random.default_rng(7)
    s0, n1, lb, nb, P = 1, 512, 8, 2, 4
    mp, s1 = lb * nb, n1 // P
    a = _cx(rng, (s0, n1, mp))
    bufs = [np.zeros(s0 * s1 * mp, complex) for _ in range(P)]
    with _env(NDFB_TRACE=1):
        be.ndfft_scatter_out(a.reshape(s0, n1, nb, lb), be.FftHandler(n1), 1, out_shape=(s0, n1, nb, lb), out_strides=(s1 * mp, lb, s1 * lb, 1),
                             out_block=s1, block_ptrs=[b.ctypes.data for b in bufs])
    assert "bulk-async copies: 4 x 16384 bytes" in capfd.readouterr().err
    fa = np.fft.fft(a, axis=1)
    for d in range(P):
        got = bufs[d].reshape(s0, nb, s1, lb).transpose(0, 2, 1, 3).reshape(s0, s1, mp)
        assert orc.rel_l2(got, fa[:, d * s1:(d + 1) * s1, :]) < 1e-12
    with _env(NDFB_NO_BULK_STORE=1):
        b2 = [np.zeros(s0 * s1 * mp, complex) for _ in range(P)]
        be.ndfft_scatter_out(a.reshape(s0, n1, nb, lb), be.FftHandler(n1), 1, out_shape=(s0, n1, nb, lb), out_strides=(s1 * mp, lb, s1 * lb, 1),
                             out_block=s1, block_ptrs=[b.ctypes.data for b in b2])
    for d in range(P):
        assert np.array_equal(b2[d], bufs[d])


def test_signal_and_wait_hints(be, capfd):
    """Producer / consumer launch hints (ndfb_hint_next_launch_signal / _wait): the r2c pass counts finished rows per plane,
    the persistent exchange pass waits per plane.  Under the emulator launches run one after the other, so every wait is
    already satisfied; the counters, the persistent tile loop and the results are what is checked."""
    rng = np.random.default_rng(8)
    s0, n1, n2 = 3, 64, 128
    m = n2 // 2 + 1
    mp = 80
    x = rng.uniform(-1, 1, (s0, n1, n2))
    a_pad = np.zeros((s0, n1, mp), complex)
    cnt = np.zeros(s0, np.uint32)
    dll = be.lib.dll
    dll.ndfb_hint_next_launch_signal(cnt.ctypes.data, n1)
    be.ndfft_r2c(x, a_pad[:, :, :m], be.R2cFftHandler(n2), 2)
    # host arrays with gaps are packed first: the counters still count every row of every plane
    assert cnt.tolist() == [n1] * s0
    P, s1 = 2, n1 // 2
    bufs = [np.zeros((s0, s1, mp), complex) for _ in range(P)]
    with _env(NDFB_TRACE=1):
        dll.ndfb_hint_next_launch_wait(cnt.ctypes.data, mp, n1, 1)
        be.ndfft_scatter_out(a_pad, be.FftHandler(n1), 1, out_shape=(s0, n1, mp), out_strides=(s1 * mp, mp, 1), out_block=s1,
                             block_ptrs=[b.ctypes.data for b in bufs])
    assert "persistent consumer launch" in capfd.readouterr().err
    want = np.fft.fft(np.fft.rfft(x, axis=2), axis=1)
    for p in range(P):
        assert orc.rel_l2(bufs[p][:, :, :m], want[:, p * s1:(p + 1) * s1, :]) < 1e-12
    # a call that cannot honour the hint says so
    dll.ndfb_hint_next_launch_wait(cnt.ctypes.data, mp, n1, 1)
    y = np.zeros((4, 17), complex)
    with pytest.raises(Exception, match="hint"):
        be.ndfft(np.zeros((4, 17), complex), y, be.FftHandler(17), 1)
