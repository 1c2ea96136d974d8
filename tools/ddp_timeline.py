"""Where does the multi-GPU step lose time?    Run under torchrun on N GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_timeline.py

For KAN-VGG16 @224, batch 64 per GPU (the bench workload) it measures, on the device and as the max over ranks:
  * ms/step with the gradient all-reduce (DDP) and without it (DDP.no_sync) -> the EXPOSED cost of the collective;
  * from a torch.profiler (CUPTI) trace of two steps on rank 0: total NCCL kernel time, the part of it that runs concurrently with
    this library's kernels, and the tail after the last compute kernel of the backward pass;
  * the same with bucket sizes 25 / 64 / 256 MB and with the bf16 gradient-compression hook.
Prints one JSON line per variant (rank 0)."""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kanconv_b200 as K  # noqa: E402
from kanconv_b200.models import vggkan  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
BATCH = int(os.environ.get("DDP_BATCH", "64"))
torch.manual_seed(0)
model = vggkan(3, 1000, arch="VGG16", classifier_type="Linear", expected_feature_shape=(7, 7), spline_order=3, grid_size=5).to(dev).train()
lossf = nn.CrossEntropyLoss()
g = torch.Generator().manual_seed(1234 + rank)
x = torch.randn(BATCH, 3, 224, 224, generator=g).to(dev)
y = torch.randint(0, 1000, (BATCH,), generator=g).to(dev)


def timed(fn, steps=8, warmup=3):
    for _ in range(warmup):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def union(iv):
    iv = sorted(iv)
    out = []
    for s, e in iv:
        if out and s <= out[-1][1]:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([s, e])
    return out


def overlap(a, b):
    a, b = union(a), union(b)
    i = j = 0
    tot = 0.0
    while i < len(a) and j < len(b):
        s, e = max(a[i][0], b[j][0]), min(a[i][1], b[j][1])
        if e > s:
            tot += e - s
        if a[i][1] < b[j][1]:
            i += 1
        else:
            j += 1
    return tot


variants = [("bucket64", 64, None), ("bucket25", 25, None), ("bucket256", 256, None), ("bucket64_bf16hook", 64, "bf16")]
for name, cap, hook in variants:
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True, bucket_cap_mb=cap)
    if hook == "bf16":
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        net.register_comm_hook(None, default_hooks.bf16_compress_hook)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)

    def step():
        opt.zero_grad(set_to_none=True)
        lossf(net(x), y).backward()
        opt.step()

    def step_nosync():
        opt.zero_grad(set_to_none=True)
        with net.no_sync():
            lossf(net(x), y).backward()
        opt.step()

    ms_sync = timed(step)
    ms_nosync = timed(step_nosync)
    rec = {"variant": name, "n_gpus": world, "per_gpu_batch": BATCH, "ms_per_step_ddp": round(ms_sync, 3), "ms_per_step_no_allreduce": round(ms_nosync, 3),
           "exposed_allreduce_ms": round(ms_sync - ms_nosync, 3), "grad_mbytes": round(sum(p.numel() for p in model.parameters()) * 4 / 1e6, 1)}
    # kernel timeline of two steps (rank 0)
    from torch.profiler import ProfilerActivity, profile
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        step()
        torch.cuda.synchronize()
    if rank == 0:
        nccl, mine, other = [], [], []
        for ev in prof.events():
            if ev.device_type != torch.autograd.DeviceType.CUDA:
                continue
            s, e = ev.time_range.start, ev.time_range.end
            nm = ev.name
            (nccl if "nccl" in nm.lower() else mine if "kc_" in nm else other).append((s, e))
        if nccl and mine:
            tn = sum(e - s for s, e in union(nccl))
            rec.update({"nccl_kernel_ms_per_step": round(tn / 2e3, 3), "nccl_concurrent_with_kc_kernels_ms_per_step": round(overlap(nccl, mine) / 2e3, 3),
                        "nccl_kernels_per_step": len(nccl) // 2, "kc_kernel_ms_per_step": round(sum(e - s for s, e in union(mine)) / 2e3, 3),
                        "nccl_tail_after_last_kc_kernel_ms": round(max(0.0, max(e for _, e in nccl) - max(e for _, e in mine)) / 1e3, 3)})
        print(json.dumps(rec), flush=True)
    del net, opt
dist.destroy_process_group()
