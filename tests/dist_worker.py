"""Worker for tests/test_dist.py: run under torchrun with backend gloo (CPU, SIMT-emulated kernels) or nccl (GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    backend = sys.argv[1]
    n0, n1, n2 = (int(v) for v in sys.argv[2:5])
    dtype = np.dtype(sys.argv[5]) if len(sys.argv) > 5 else np.dtype(np.float64)
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=device)
        be = None
    else:
        device = torch.device("cpu")
        dist.init_process_group("gloo")
        from emu_backend import emu_backend
        be = emu_backend()
    from ndrustfft_b200.dist import SlabR2cFft3d, shard_bounds, sharded_apply
    import ndrustfft_b200 as nb
    rank, world = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(1234)
    xg = rng.uniform(-1, 1, (n0, n1, n2)).astype(dtype)        # same global array on every rank
    want = np.fft.rfftn(xg.astype(np.float64), axes=(0, 1, 2))  # = fft(axis0) . fft(axis1) . rfft(axis2)
    plan = SlabR2cFft3d((n0, n1, n2), dtype, device=device, backend=be, chunks=int(os.environ.get('NDFB_TEST_CHUNKS', '2')),
                        peer={'auto': 'auto', 'on': True, 'off': False}[os.environ.get('NDFB_TEST_PEER', 'auto')])
    lo, hi = shard_bounds(n0, world, rank)
    x = torch.from_numpy(xg[lo:hi].copy()).to(device)
    X = plan.forward(x)
    c0, c1 = shard_bounds(n1, world, rank)
    got = X.cpu().numpy()
    ref = want[:, c0:c1, :]
    tol = 1e-12 if dtype == np.float64 else 1e-5
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert err < tol, f"rank {rank}: forward rel L2 {err}"
    back = plan.inverse(X)
    err_b = np.linalg.norm(back.cpu().numpy() - xg[lo:hi]) / np.linalg.norm(xg[lo:hi])
    assert err_b < tol, f"rank {rank}: round trip rel L2 {err_b}"
    # lane-sharded single-axis call: no collective, every rank its slice
    backend_obj = be or nb._default_backend()
    xc = (rng.uniform(-1, 1, (n0, n1)) + 1j * rng.uniform(-1, 1, (n0, n1))).astype(np.complex64 if dtype == np.float32 else np.complex128)
    xin = torch.from_numpy(xc).to(device)
    out = torch.zeros_like(xin)
    h = backend_obj.FftHandler(n1, dtype, device.index or 0 if device.type == "cuda" else 0)
    lo2, hi2 = sharded_apply(backend_obj.ndfft, xin, out, h, 1, 0, world, rank)
    refc = np.fft.fft(xc.astype(np.complex128), axis=1)
    e2 = np.linalg.norm(out.cpu().numpy()[lo2:hi2] - refc[lo2:hi2]) / np.linalg.norm(refc[lo2:hi2])
    assert e2 < tol, e2
    assert not out.cpu().numpy()[:lo2].any() and not out.cpu().numpy()[hi2:].any()
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK peer={plan.peer} world={world} fwd={err:.2e} back={err_b:.2e} shard={e2:.2e} sent_per_rank={plan.bytes_sent_per_rank()}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
