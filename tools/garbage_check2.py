"""Garbage dependence at model level: prefixes of KAN-VGG16.features (64x64 input, batch 4, bf16), loss = <out, fixed random>:
which prefix length first gives parameter gradients that depend on the contents of free GPU memory?"""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
from kanconv_b200.models.kan_vgg import vggkan
dev = torch.device("cuda")
K.set_precision("bf16")
torch.manual_seed(0)
model = vggkan(3, 1000, arch="VGG16", classifier_type="Linear", expected_feature_shape=(7, 7), spline_order=3, grid_size=5).to(dev)
model.eval()
feats = list(model.features.children())
print([type(f).__name__ for f in feats])
x = torch.randn(4, 3, 64, 64, device=dev)
def run(k, val, train_mode):
    torch.cuda.synchronize()
    big = torch.full((2 * 1024 ** 3 // 4,), val, device=dev, dtype=torch.float32)
    small = [torch.full((s,), val, device=dev, dtype=torch.float32) for s in (16, 256, 4096, 65536, 200000) for _ in range(8)]
    del big, small
    model.zero_grad(set_to_none=True)
    h = x
    for f in feats[:k]:
        h = f(h)
    torch.manual_seed(1)
    w = torch.randn_like(h)
    (h * w).sum().backward()
    torch.cuda.synchronize()
    return {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
for k in range(1, len(feats) + 1):
    a = run(k, 0.0, False); b = run(k, 3000.0, False)
    diff = [f"{n} ({float((a[n] - b[n]).abs().max() / a[n].abs().max().clamp_min(1e-30)):.1e})" for n in a if not torch.equal(a[n], b[n])]
    print(f"prefix {k:2d} (... {type(feats[k - 1]).__name__}): " + ("identical" if not diff else f"{len(diff)} DIFFERENT: " + ", ".join(diff[:4])), flush=True)
    if len(diff) > 0 and k > 6:
        break
