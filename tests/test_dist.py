"""Multi-process tests of the slab-decomposed 3-D transform and lane sharding.
CPU: world_size 2 over gloo with the SIMT-emulated kernels.  GPU (-m gpu): the same worker over NCCL when >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "dist_worker.py")


def _run(backend, nproc, shape, dtype, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, backend, *map(str, shape), dtype]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "DIST_OK" in p.stdout, p.stdout[-2000:]
    return p.stdout


def test_slab_fft3d_gloo_world2():
    from emu_backend import emu_backend
    emu_backend()                                   # build the emulation library once, before the ranks race for it
    out = _run("gloo", 2, (8, 12, 10), "float64", 29631)
    assert "world=2" in out


def test_slab_fft3d_gloo_world4_f32():
    from emu_backend import emu_backend
    emu_backend()
    _run("gloo", 4, (8, 8, 16), "float32", 29632)


def test_shard_bounds():
    from ndrustfft_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 513):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


@pytest.mark.gpu
def test_slab_fft3d_nccl():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    out = _run("nccl", 2, (64, 64, 64), "float64", 29641)           # default: fused peer stores when symmetric memory works
    os.environ["NDFB_TEST_PEER"] = "off"
    try:
        out2 = _run("nccl", 2, (64, 64, 64), "float64", 29642)      # NCCL all_to_all_single baseline path
    finally:
        del os.environ["NDFB_TEST_PEER"]
    assert "peer=False" in out2
    print(out, out2)
