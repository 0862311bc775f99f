import os, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = {}
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1024, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024,), torch.float32)
    ok["symm"] = float(peer[0].item())
    ok["ptrs"] = [hex(p) for p in hdl.buffer_ptrs]
except Exception as e:
    ok["symm_err"] = repr(e)[:300]
try:
    x = torch.full((1024,), float(rank + 10), device=dev)
    st = x.untyped_storage()
    info = st._share_cuda_()
    objs = [None] * world
    dist.all_gather_object(objs, info)
    pinfo = objs[(rank + 1) % world]
    pst = torch.UntypedStorage._new_shared_cuda(*pinfo)
    pt = torch.tensor([], dtype=torch.float32, device=pst.device).set_(pst, 0, (1024,))
    ok["ipc"] = float(pt[0].item()); ok["ipc_dev"] = str(pt.device); ok["ipc_ptr"] = hex(pt.data_ptr())
except Exception as e:
    ok["ipc_err"] = repr(e)[:300]
print("PROBE", rank, ok, flush=True)
dist.barrier()
dist.destroy_process_group()
