#!/usr/bin/env python
"""bench.py - headline benchmark of the KAN-convolution hot path: KAN-VGG training images/s on synthetic data.

    python bench.py --gpus N --steps K --warmup W                      # this repo's CUDA path ("b200" arm)
    python bench.py --impl reference --gpus N --steps K --warmup W     # the reference algorithm on the host CPU cores

Metric (BASELINE.json): "KAN-VGG train images/sec" - one step = forward + CrossEntropy + backward + AdamW on one batch of
synthetic images.  Default workload = BASELINE config 5: KAN-VGG16 (KANConv2D, spline_order 3, grid_size 5, SiLU base
activation, InstanceNorm, Linear head, expected_feature_shape (7,7)) on 3x224x224, per-GPU batch 64, weak scaling.
Prints ONE JSON line (rank 0).  For N > 1 launch with torch.distributed.run (one rank per GPU, NCCL).

Keys beyond the base contract:
  roofline      - dominant kernel of the step (by summed device time), measured live with CUDA events on the launching
                  stream in a separate instrumented pass of the SAME step: achieved = algorithmic FLOPs per launch
                  (2*N*Ho*Wo*Cout*Cin*(nb+1)*kh*kw) / mean launch duration; peak = MEASURED_PEAKS.json.
  cpu_baseline  - the reference's own modules when its tree is importable (/root/reference in the dev container, or
                  baseline/_ref; kind "reference"), else the oracle port of the same algorithm (oracle/kan_oracle.py; kind
                  "port" - the reference is a Python project and does not travel to the GPU box), PyTorch CPU, all host cores,
                  timed on a bounded sample of the same workload (rank 0, N = 1 only).
  gpu_eager_baseline - informational: the same eager-PyTorch modules (reference or port) on THIS GPU, cuDNN/cuBLAS with
                  TF32 off and on, at the largest power-of-two batch that fits - what a user of the reference runs today.
  ddp_selfcheck - N > 1 only: max relative deviation of the DDP-averaged gradients of a small sharded batch from the
                  single-process gradients of the concatenated batch (outside the timed region).
  e2e           - same metric through the public nn.Module API with HOST inputs: every step copies its pinned-host batch
                  to the device and reads the loss back.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

TRAFFIC_FILE = "r2_traffic_b64.json"      # ncu DRAM-traffic capture of this workload (tools/ncu_round.sh + summarize_profiles.py)

WORKLOADS = {
    # name: (arch, H=W, classes, default per-GPU batch, expected_feature_shape, cpu sample batch)
    "kan_vgg16_224": ("VGG16", 224, 1000, 64, (7, 7), 1),
    "kan_vgg11_32": ("VGG11", 32, 10, 512, (1, 1), 16),
    "kan_vgg16_small_32": ("VGG16_small", 32, 10, 256, (1, 1), 32),
}


def vgg_conv_shapes(arch, hw, cin=3):
    """[(cin, cout, h)] of the KAN conv layers of a KAN-VGG (models/kan_vgg.py:119-130 semantics)."""
    from kanconv_b200.models import cfgs as VGG_CFGS
    out, c, h = [], cin, hw
    for v in VGG_CFGS[arch]:
        if v == "M":
            h //= 2
        else:
            out.append((c, int(v), h))
            c = int(v)
    return out


def step_flops(arch, hw, batch):
    """Dense-equivalent FLOPs of one training step: fwd F, bwd 2F, no dgrad for the first layer."""
    total = 0.0
    for i, (ci, co, h) in enumerate(vgg_conv_shapes(arch, hw)):
        f = 2.0 * batch * h * h * co * ci * 9 * 9
        total += f * (3 if i > 0 else 2)
    return total


class ClockSampler:
    """nvidia-smi sampling of SM clock / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            p = [v.strip() for v in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), float(p["hbm_gbs"]), "measured"
    except Exception:
        return 1400.0, 1590.0, 6650.0, "fallback"


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------------------------------
def reference_model(workload):
    """-> (nn.Module on the CPU, kind): the reference's own vggkan() when its tree is importable, else the oracle port."""
    arch, hw, classes, _, feat, _ = WORKLOADS[workload]
    for cand in (os.environ.get("KAN_REFERENCE"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if not cand or not os.path.isfile(os.path.join(cand, "models", "kan_vgg.py")):
            continue
        try:
            import contextlib
            import io
            sys.dont_write_bytecode = True
            sys.path.insert(0, cand)
            import models.kan_vgg as ref_vgg
            if "VGG11" not in ref_vgg.cfgs:       # BASELINE config 3: absent upstream, injected as in the survey
                ref_vgg.cfgs["VGG11"] = [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512]
            torch.manual_seed(0)
            with contextlib.redirect_stdout(io.StringIO()):
                m = ref_vgg.vggkan(3, classes, arch=arch, classifier_type="Linear", expected_feature_shape=feat, spline_order=3,
                                   grid_size=5)
            return m, "reference"
        except Exception as e:      # incomplete tree / missing dependency: fall back to the port, and say so
            print(f"bench.py: reference at {cand} not usable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        finally:
            if cand in sys.path:
                sys.path.remove(cand)
    from oracle import kan_oracle as O
    torch.manual_seed(0)
    return O.OracleVGG(3, classes, arch=arch, expected_feature_shape=feat, dropout_linear=0.5), "port"


def cpu_reference_run(workload, steps, warmup, sample_batch=None):
    arch, hw, classes, _, feat, cpu_b = WORKLOADS[workload]
    b = sample_batch or cpu_b
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, kind = reference_model(workload)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
    lossf = nn.CrossEntropyLoss()
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(b, 3, hw, hw, generator=g)
    y = torch.randint(0, classes, (b,), generator=g)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = lossf(model(x), y)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return {"value": b / mean, "unit": "images/s", "cores": cores, "kind": kind,
            "sample": f"{steps} step(s) of batch {b} after {warmup} warm-up ({workload}, fp32, torch {torch.__version__} CPU, "
                      f"{torch.get_num_threads()} threads; min {min(times) * 1e3:.0f} ms, mean {mean * 1e3:.0f} ms per step)",
            "ms_per_step": mean * 1e3, "batch": b}


def gpu_eager_baseline(workload, dev):
    """Informational: the eager-PyTorch modules of the reference (or the port) on this GPU.  TF32 off and on; batch doubled
    while the measured peak memory says the next size still fits in 70 % of the device."""
    arch, hw, classes, _, feat, _ = WORKLOADS[workload]
    out = {"unit": "images/s", "note": "eager PyTorch (cuDNN / cuBLAS) forward + CrossEntropy + backward + AdamW of the reference "
                                       "modules on the same GPU; materialises the expanded basis tensor"}
    try:
        model, kind = reference_model(workload)
        model = model.to(dev).train()
        out["kind"] = kind
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4)
        lossf = nn.CrossEntropyLoss()
        total = torch.cuda.get_device_properties(dev).total_memory
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)

        def run(b, tf32, steps=2):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            g = torch.Generator().manual_seed(1234)
            x = torch.randn(b, 3, hw, hw, generator=g).to(dev)
            y = torch.randint(0, classes, (b,), generator=g).to(dev)
            ts = []
            for i in range(1 + steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                opt.zero_grad(set_to_none=True)
                lossf(model(x), y).backward()
                opt.step()
                e1.record()
                torch.cuda.synchronize()
                if i > 0:
                    ts.append(e0.elapsed_time(e1))
            return b / (min(ts) * 1e-3)

        try:
            b, best = 1, {}
            while True:
                torch.cuda.reset_peak_memory_stats(dev)
                best = {"batch": b, "tf32_off": run(b, False), "tf32_on": run(b, True)}
                peak = torch.cuda.max_memory_allocated(dev)
                out.update(best, peak_gb=round(peak / 2 ** 30, 1))
                if b >= 64 or peak * 2.2 > 0.7 * total:
                    break
                b *= 2
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    except Exception as e:     # informational leg: never let it take the bench line down
        out["error"] = f"{type(e).__name__}: {e}"[:200]
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arch, hw, classes, batch, feat, _ = WORKLOADS[args.workload]
    r = cpu_reference_run(args.workload, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "KAN-VGG train images/sec", "value": r["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "arch": arch, "image": [3, hw, hw], "classes": classes,
                       "per_gpu_batch": batch, "sampled_batch": r["batch"],
                       "impl": ("the reference's own modules (CPU)" if r["kind"] == "reference"
                                else "oracle port of the reference (CPU; the Python reference tree is not on this box)")},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# b200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import kanconv_b200 as K
    from kanconv_b200 import functional as KF
    from kanconv_b200.models import vggkan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (b200 arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = K._lib.load()
    K.set_precision(args.precision)

    arch, hw, classes, def_batch, feat, _ = WORKLOADS[args.workload]
    batch = args.batch or def_batch
    torch.manual_seed(0)
    model = vggkan(3, classes, arch=arch, classifier_type="Linear", expected_feature_shape=feat, spline_order=3,
                   grid_size=5).to(dev)
    model.train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True,
                                                        bucket_cap_mb=64)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
    lossf = nn.CrossEntropyLoss()

    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(batch, 3, hw, hw, generator=g).pin_memory()
    y_host = torch.randint(0, classes, (batch,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        loss = lossf(net(x), y)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg -------------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step(x_dev, y_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.kc_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    host_t = []
    for _ in range(args.steps):
        t_h = time.perf_counter()
        step(x_dev, y_dev)
        host_t.append(time.perf_counter() - t_h)
    e1.record()
    barrier()
    host_ms = sorted(host_t)[len(host_t) // 2] * 1e3      # median host time to ENQUEUE one step (no synchronisation inside)
    launches = lib.kc_launch_count() - l0
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end leg: host inputs, H2D per step, loss read back per step ----------------------------------------
    for _ in range(min(2, args.warmup)):
        step(x_host.to(dev, non_blocking=True), y_host.to(dev, non_blocking=True)).item()
    f0 = torch.cuda.Event(enable_timing=True)
    f1 = torch.cuda.Event(enable_timing=True)
    # Every step's batch is copied from pinned host memory inside the timed region and every step's loss is copied back to
    # pinned host memory.  Like a prefetching data loader, the copy of step i+1 runs on a second stream while step i
    # computes (two device buffers), and the loss read-back is asynchronous (checked when the loop ends), so the GPU never
    # waits for the host between steps.
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(x_dev), torch.empty_like(y_dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(consumed[i % 2])      # step i-2 has finished reading this buffer
            bufs[i % 2][0].copy_(x_host, non_blocking=True)
            bufs[i % 2][1].copy_(y_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    barrier()
    t0 = time.perf_counter()
    f0.record()
    prefetch(0)
    for i in range(args.steps):
        torch.cuda.current_stream().wait_event(ready[i % 2])
        loss = step(*bufs[i % 2])
        consumed[i % 2].record()
        loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)      # device -> host read of the loss
        if i + 1 < args.steps:
            prefetch(i + 1)
    f1.record()
    barrier()
    ms_e2e = max_over_ranks(f0.elapsed_time(f1)) / args.steps
    wall_e2e = (time.perf_counter() - t0) / args.steps * 1e3
    last = float(loss_host[-1])

    # ---- roofline of the dominant kernel (instrumented pass of the same step, CUDA events per library call) --------
    KF.profile_begin()
    step(x_dev, y_dev)
    torch.cuda.synchronize()
    prof = KF.profile_end()
    sust, burst, hbm, peak_src = measured_peaks()
    roof = None
    if prof:
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        name, st = top
        total_ms = sum(v["ms"] for v in prof.values())
        if st["flops"] > 0:
            ach = st["flops"] / (st["ms"] * 1e-3) / 1e12
            roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": sust, "unit": "TFLOP/s", "frac": ach / sust,
                    "traffic": None, "launches_per_step": st["calls"], "ms_per_launch": st["ms"] / st["calls"],
                    "share_of_library_time": st["ms"] / total_ms, "peak_source": peak_src + " (bf16_tflops_sustained; burst %.0f)" % burst}
        else:
            ach = st["bytes"] / (st["ms"] * 1e-3) / 1e9
            roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                    "traffic": None, "launches_per_step": st["calls"], "ms_per_launch": st["ms"] / st["calls"],
                    "share_of_library_time": st["ms"] / total_ms, "peak_source": peak_src}
        roof["by_kernel_ms"] = {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        # DRAM bytes per launch of that kernel from the committed ncu capture of this workload (profiles/).  The capture
        # records the hash of the CUDA sources it was taken with; a capture of other kernels is refused (traffic = null).
        tpath = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
        if args.workload == "kan_vgg16_224" and batch == 64 and os.path.exists(tpath):
            src_hash = K._lib.source_hash()
            with open(tpath) as fh:
                cap = json.load(fh)
            if cap.get("source_hash") != src_hash:
                roof["traffic_stale"] = (f"profiles/{TRAFFIC_FILE} was captured with other kernel sources "
                                         f"({str(cap.get('source_hash'))[:12]} != {src_hash[:12]}); not used")
            else:
                kern = cap["kernels"]
                # template instantiations of one kernel are listed separately in the capture (kc_wgrad_tc_kernel<64>, <128>)
                hits = [v for k, v in kern.items() if k == name or k.startswith(name + "<")]
                if hits:
                    nl = sum(v["launches_per_step"] for v in hits)
                    tot = sum(v["dram_read_bytes_per_step"] + v["dram_write_bytes_per_step"] for v in hits)
                    roof["traffic"] = int(tot / max(nl, 1e-9))
                    roof["traffic_source"] = (f"profiles/{TRAFFIC_FILE} (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean per "
                                              f"launch; sources {src_hash[:12]})")

    # ---- N > 1: gradient equality of the sharded step (outside the timed region) -----------------------------------
    selfcheck = None
    if world > 1:
        # At the bench's own resolution: on 64x64 images the tail of KAN-VGG16 runs InstanceNorm over 4x4 planes, whose backward
        # amplifies a 1e-7 perturbation (torch's head computes the gradient of a batch of 2 and of 2 * world images in a different
        # summation order) ten-fold per layer - 3e-3..5e-3 at the first layer although every kernel of this library gives
        # bit-identical per-sample results in both batches (tools/shard_check.py, tests/test_ddp_gpu.py).
        per, side = 2, hw
        gs = torch.Generator().manual_seed(4321)
        xa = torch.randn(per * world, 3, side, side, generator=gs).to(dev)
        ya = torch.randint(0, classes, (per * world,), generator=gs).to(dev)
        model.eval()                      # head Dropout off: the two runs must see the same function
        opt.zero_grad(set_to_none=True)
        lossf(net(xa[rank * per:(rank + 1) * per]), ya[rank * per:(rank + 1) * per]).backward()      # DDP: all-reduced mean
        torch.cuda.synchronize()
        got = [p.grad.detach().clone() for p in model.parameters()]
        opt.zero_grad(set_to_none=True)
        lossf(model(xa), ya).backward()                                                             # one process, whole batch
        worst, num, den = 0.0, 0.0, 0.0
        for a_, p in zip(got, model.parameters()):
            num += float((a_ - p.grad).double().square().sum())
            den += float(p.grad.double().square().sum())
            if p.numel() > 1:      # PReLU slopes are scalar sums with heavy cancellation: in the global L2 only
                worst = max(worst, float((a_ - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-30)))
        worst = max_over_ranks(worst)
        l2 = max_over_ranks((num / max(den, 1e-300)) ** 0.5)
        opt.zero_grad(set_to_none=True)
        model.train()
        selfcheck = {"grad_rel_l2": l2, "grad_max_rel": worst, "images": per * world, "image": [3, side, side],
                     "what": "DDP-averaged gradients of a sharded batch vs one process on the whole batch: relative L2 over all "
                             "parameters; max over weight tensors and ranks of max|g_ddp - g_single| / max|g_single| (same kernels, "
                             "bit-identical per-sample results; the order of the cross-sample fp32 sums differs, and the rounding of torch's "
                             "head differs between the two batch sizes, which the normalised layers amplify on the way back)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(args.workload, 3, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    eager = None
    if world == 1 and not args.no_gpu_eager_baseline:
        del net, opt
        model.to("cpu")
        KF.clear_pack_cache()
        torch.cuda.empty_cache()
        eager = gpu_eager_baseline(args.workload, dev)
    gb = batch * world
    flops = step_flops(arch, hw, batch)
    line = {
        "metric": "KAN-VGG train images/sec", "value": gb / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
        "config": {"workload": args.workload, "arch": arch, "image": [3, hw, hw], "classes": classes, "per_gpu_batch": batch,
                   "global_batch": gb, "parallelism": f"dp{world}", "optimizer": "AdamW(fused)", "precision": args.precision,
                   "l2": "inputs+activations per step >> L2 (no flush needed)",
                   "step_tflops_dense_equiv": flops / 1e12, "achieved_tflops_per_gpu": flops / (ms * 1e-3) / 1e12},
        "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(x_host.numel() * 4 + y_host.numel() * 8),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "wall_ms_per_step": wall_e2e, "last_loss": last},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(host_ms, 2), "clocks": clocks, "roofline": roof,
        "cpu_baseline": cpu,
    }
    if eager is not None:
        line["gpu_eager_baseline"] = eager
    if selfcheck is not None:
        line["ddp_selfcheck"] = selfcheck
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="kan_vgg16_224", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--precision", default="auto", choices=["auto", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "b200" and args.gpus != world:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torch.distributed.run with --nproc-per-node {args.gpus} (WORLD_SIZE={world})")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
