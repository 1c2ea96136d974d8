"""BASELINE configs 2-4 at their full sizes (parity-test cases, not the bench line): a few training steps each on one GPU,
one JSON record per config (ms/step, images/s, dense-equivalent TFLOP/s or algorithmic GB/s against the measured peaks, per-kernel
time split, all gradients finite).  python tools/other_configs.py > gpurun_out/r2/other_configs.jsonl  (kept under profiles/)."""
import json
import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kanconv_b200 as K  # noqa: E402
from kanconv_b200.models import kan_mobilenetv2, kan_vgg  # noqa: E402

try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def step_time(model, x, y, steps=5):
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    lossf = nn.CrossEntropyLoss()

    def one():
        opt.zero_grad(set_to_none=True)
        loss = lossf(model(x), y)
        loss.backward()
        opt.step()
        return loss
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    ok = all(torch.isfinite(p.grad).all().item() for p in model.parameters() if p.grad is not None)
    K.functional.profile_begin()
    one()
    prof = K.functional.profile_end()
    kern = {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]}
    flops = sum(v["flops"] for v in prof.values())          # dense-equivalent, SURVEY 8(d), fwd + dgrad + wgrad as launched
    return e0.elapsed_time(e1) / steps, float(loss.detach()), ok, kern, flops


def record(config, what, batch, ms, loss, ok, kern, flops, extra=None):
    r = {"config": config, "workload": what, "batch": batch, "ms_per_step": round(ms, 3), "images_per_s": round(batch / ms * 1e3, 1),
         "loss": round(loss, 4), "finite_grads": ok, "dense_equiv_tflops": round(flops / ms / 1e9, 1),
         "frac_of_sustained_bf16_peak": round(flops / ms / 1e9 / PEAKS["bf16_tflops_sustained"], 3), "kernel_ms": kern}
    r.update(extra or {})
    print(json.dumps(r), flush=True)


torch.manual_seed(0)
dev = "cuda"
# config 2: ChebyKAN / GRAMKAN degree-3 stack L(64,128,3,p=1) -> L(128,128,3,p=1), batch 256 x 64 x 32 x 32 (a small
# head turns the stack into a training step)
for name in ("ChebyKAN", "GRAMKAN"):
    f = K.CONV_KAN_FACTORY[name]
    m = nn.Sequential(f(64, 128, 3, padding=1), f(128, 128, 3, padding=1), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(128, 10)).to(dev)
    x = torch.randn(256, 64, 32, 32, device=dev)
    y = torch.randint(0, 10, (256,), device=dev)
    record(2, f"{name} stack 64->128->128, degree 3, 256x64x32x32", 256, *step_time(m, x, y))
# config 3: KAN-VGG11 on CIFAR-shaped input, batch 512 (tail maps 2x2, K = 41 472)
m = kan_vgg.vggkan(3, 10, arch="VGG11", classifier_type="Linear", expected_feature_shape=(1, 1), spline_order=3, grid_size=5).to(dev)
x = torch.randn(512, 3, 32, 32, device=dev)
y = torch.randint(0, 10, (512,), device=dev)
record(3, "KAN-VGG11 (KANConv2D, spline_order 3, grid_size 5), 512x3x32x32", 512, *step_time(m, x, y))
del m
# config 4: FastKAN MobileNetV2 on ImageNet-shaped input, batch 128: HBM-bound (~160 MB / image of algorithmic
# activation traffic forward + backward, fp32)
m = kan_mobilenetv2.mobilenet_v2_kan(num_classes=1000, kan_conv="FastKAN", classifier_type="Linear").to(dev)
x = torch.randn(128, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (128,), device=dev)
ms, loss, ok, kern, flops = step_time(m, x, y, steps=3)
gbs = 160e6 * 128 / (ms * 1e-3) / 1e9
record(4, "FastKAN MobileNetV2 (RBF, grid_size 5), 128x3x224x224", 128, ms, loss, ok, kern, flops,
       {"algorithmic_gb_per_s": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAKS["hbm_gbs"], 3),
        "algorithmic_bytes_per_image": 160e6})
