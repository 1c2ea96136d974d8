"""Do results depend on what free GPU memory contains?  For single KAN layers of the KAN-VGG16 shapes: the same forward +
backward after the caching allocator's free blocks were filled with 0 and with 3000.0; every result must be bit-identical."""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
dev = torch.device("cuda")
K.set_precision("bf16")
def run(m, x, gy, val):
    torch.cuda.synchronize()
    big = torch.full((2 * 1024 ** 3 // 4,), val, device=dev, dtype=torch.float32)
    small = [torch.full((s,), val, device=dev, dtype=torch.float32) for s in (16, 256, 4096, 65536, 200000) for _ in range(8)]
    del big, small
    m.zero_grad(set_to_none=True)
    xx = x.clone().requires_grad_(True)
    y = m(xx)
    y.backward(gy)
    torch.cuda.synchronize()
    out = {"y": y.detach().clone(), "dx": xx.grad.clone()}
    out.update({"d " + k: p.grad.clone() for k, p in m.named_parameters()})
    return out
bad = 0
for (cin, cout, hw, n) in [(3, 64, 64, 4), (64, 64, 64, 4), (64, 128, 32, 4), (128, 128, 32, 4), (128, 256, 16, 4), (256, 256, 16, 4),
                           (256, 512, 8, 4), (512, 512, 8, 4), (512, 512, 4, 4), (64, 64, 224, 2), (128, 128, 112, 2)]:
    torch.manual_seed(0)
    m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU).to(dev)
    x = torch.randn(n, cin, hw, hw, device=dev)
    gy = torch.randn(n, cout, hw, hw, device=dev)
    a = run(m, x, gy, 0.0); b = run(m, x, gy, 3000.0); c = run(m, x, gy, 0.0)
    diff = [f"{k} ({float((a[k] - b[k]).abs().max() / a[k].abs().max()):.1e})" for k in a if not torch.equal(a[k], b[k])]
    diff2 = [k for k in a if not torch.equal(a[k], c[k])]
    bad += len(diff)
    print(f"{cin:4d} -> {cout:4d} @ {hw:3d} x{n}: 0 vs 3000: " + ("identical" if not diff else "DIFFERENT " + ", ".join(diff)) +
          " | 0 vs 0 again: " + ("identical" if not diff2 else "DIFFERENT " + ", ".join(diff2)), flush=True)
sys.exit(1 if bad else 0)
