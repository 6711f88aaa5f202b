"""Does this box support torch symmetric memory with NVLS multicast?  torchrun --nproc-per-node N tools/symm_probe.py"""
import os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from torch._C._distributed_c10d import _SymmetricMemory
    try:
        print(rank, "has_multicast_support", _SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA if hasattr(torch._C._autograd, "DeviceType") else dev.type, local), flush=True)
    except Exception as e:
        print(rank, "has_multicast_support query failed:", repr(e)[:200], flush=True)
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(rank + 1.0)
    h = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
    print(rank, "rendezvous ok: multicast_ptr", hex(h.multicast_ptr) if h.multicast_ptr else h.multicast_ptr, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs],
          "signal_pad", [hex(p) for p in h.signal_pad_ptrs], "signal_pad_size", h.signal_pad_size, flush=True)
    h.barrier(0)
    try:
        torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(rank, "multimem_all_reduce_ ->", float(t[0]), float(t[-1]), "expected", world * (world + 1) / 2, flush=True)
    except Exception as e:
        print(rank, "multimem_all_reduce_ failed:", repr(e)[:300], flush=True)
    # timing of the one-shot / two-shot ops torch ships, 12 MB
    n = 3 << 20
    u = symm_mem.empty(n, dtype=torch.float32, device=dev); u.fill_(1.0)
    symm_mem.rendezvous(u, dist.group.WORLD.group_name)
    for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            for _ in range(3):
                op(u, "sum", dist.group.WORLD.group_name)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                op(u, "sum", dist.group.WORLD.group_name)
            e1.record(); torch.cuda.synchronize()
            if rank == 0:
                print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us for {n * 4 / 1e6:.1f} MB", flush=True)
            u.fill_(1.0)
        except Exception as e:
            if rank == 0:
                print(name, "failed:", repr(e)[:200], flush=True)
    v = torch.ones(n, device=dev)
    for _ in range(3):
        dist.all_reduce(v)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dist.all_reduce(v)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"nccl all_reduce: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us for {n * 4 / 1e6:.1f} MB", flush=True)
    dist.barrier()
    dist.destroy_process_group()

if __name__ == "__main__":
    main()
