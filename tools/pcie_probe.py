"""GPU: host<->device copy bandwidth per rank with ALL ranks of a torchrun launch copying at once (pinned
buffers of the size a rank moves per end-to-end evaluation at the north-star shape: 16^3 k x 128 x 128 complex
/ world each way).  Explains the e2e scaling of bench.py: if the aggregate over the ranks saturates, the
end-to-end figure of a k-sharded evaluation is bound by the host link, not by the GPUs.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 16 * 4096 * 128 * 128 // world
h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
res = {}
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = {"ms_per_copy_max_over_ranks": 1e3 * t.item(), "gbs_per_rank": nbytes / t.item() / 1e9,
                 "gbs_aggregate": world * nbytes / t.item() / 1e9}
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_rank": nbytes, **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
