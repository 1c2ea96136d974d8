"""BASELINE configs 2-4 at their full sizes (parity-test cases, not bench lines): run one training step each on the GPU, report the
time, and check that every gradient is finite.  python tools/other_configs.py"""
import os, sys, time
import torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
from kanconv_b200.models import kan_vgg, kan_mobilenetv2

def step_time(model, x, y, steps=3):
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    lossf = nn.CrossEntropyLoss()
    def one():
        opt.zero_grad(set_to_none=True)
        loss = lossf(model(x), y)
        loss.backward()
        opt.step()
        return loss
    one(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): loss = one()
    e1.record(); torch.cuda.synchronize()
    ok = all(torch.isfinite(p.grad).all().item() for p in model.parameters() if p.grad is not None)
    return e0.elapsed_time(e1) / steps, float(loss.detach()), ok

torch.manual_seed(0)
dev = "cuda"
# config 2: ChebyKAN / GRAMKAN degree-3 stack 64 -> 128, batch 256 x 32 x 32
for name in ("ChebyKAN", "GRAMKAN"):
    f = K.CONV_KAN_FACTORY[name]
    # L(64,128,3,padding=1) -> L(128,128,3,padding=1) on 256 x 64 x 32 x 32 (a small head turns it into a step)
    m = nn.Sequential(f(64, 128, 3, padding=1), f(128, 128, 3, padding=1), nn.AdaptiveAvgPool2d(1), nn.Flatten(), nn.Linear(128, 10)).to(dev)
    x = torch.randn(256, 64, 32, 32, device=dev); y = torch.randint(0, 10, (256,), device=dev)
    ms, loss, ok = step_time(m, x, y)
    print(f"config 2 {name} stack 64->128->128, batch 256x64x32x32: {ms:.2f} ms/step ({256 / ms * 1e3:.0f} images/s), loss {loss:.3f}, finite grads {ok}")
# config 3: KAN-VGG11 on CIFAR-shaped input, batch 512
m = kan_vgg.vggkan(3, 10, arch="VGG11", classifier_type="Linear", expected_feature_shape=(1, 1), spline_order=3, grid_size=5).to(dev)
x = torch.randn(512, 3, 32, 32, device=dev); y = torch.randint(0, 10, (512,), device=dev)
ms, loss, ok = step_time(m, x, y)
print(f"config 3 KAN-VGG11 batch 512x3x32x32: {ms:.2f} ms/step ({512 / ms * 1e3:.0f} images/s), loss {loss:.3f}, finite grads {ok}")
# config 4: FastKAN MobileNetV2 on ImageNet-shaped input, batch 128
for bsz in (128,):
    try:
        m = kan_mobilenetv2.mobilenet_v2_kan(num_classes=1000, kan_conv="FastKAN", kan_classifier="FastKAN").to(dev)
        x = torch.randn(bsz, 3, 224, 224, device=dev); y = torch.randint(0, 1000, (bsz,), device=dev)
        ms, loss, ok = step_time(m, x, y, steps=2)
        print(f"config 4 FastKAN MobileNetV2 batch {bsz}x3x224x224: {ms:.1f} ms/step ({bsz / ms * 1e3:.0f} images/s), loss {loss:.3f}, finite grads {ok}")
        K.functional.profile_begin()
        step_time(m, x, y, steps=1)
        prof = K.functional.profile_end()
        print("   kernel ms (2 steps):", {k: round(v["ms"], 1) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:8]})
    except Exception as e:
        print("config 4 failed:", type(e).__name__, str(e)[:300])
