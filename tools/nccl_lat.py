"""dev: latency of tiny NCCL collectives through torch.distributed (one process per GPU)."""
import os, time, torch, torch.distributed as dist
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w = dist.get_world_size()
x = torch.ones(96, dtype=torch.float64, device="cuda"); out = torch.empty(96 * w, dtype=torch.float64, device="cuda")
r = torch.ones(4, dtype=torch.float64, device="cuda")
for name, fn in (("all_gather 96 doubles", lambda: dist.all_gather_into_tensor(out, x)), ("all_reduce 4 doubles", lambda: dist.all_reduce(r))):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): fn()
    e1.record(); torch.cuda.synchronize()
    if dist.get_rank() == 0: print(f"world {w}: {name}: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per call (back to back on one stream)", flush=True)
if dist.get_rank() == 0: print("nproc", os.cpu_count(), flush=True)
dist.destroy_process_group()
