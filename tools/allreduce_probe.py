"""Multi-GPU probe: all-reduce latency of the gradient group sizes of the benchmark config (run under torchrun)."""
import os, time, torch, torch.distributed as dist
r = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
for mb in (0.5, 2, 6.3, 10.2, 20.5, 39):
    t = torch.ones(int(mb * 1e6 / 4), device="cuda")
    for _ in range(5): dist.all_reduce(t)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dist.all_reduce(t)
    e1.record(); torch.cuda.synchronize()
    if r == 0: print(f"allreduce {mb:5.1f} MB: {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us", flush=True)
dist.destroy_process_group()
