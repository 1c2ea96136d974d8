"""Hunt for reads of uninitialised memory (compute-sanitizer's initcheck is not available on the GPU pool): every buffer the
binding allocates for the library (functional._ALLOC) is pre-filled with the byte pattern C0 7F, which is NaN as bf16 and as
fp32, then single KAN layers of the KAN-VGG16 shapes run forward + backward and every result is checked for NaN.  A NaN means
some kernel consumed bytes that no kernel had written."""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
from kanconv_b200 import functional as KF
dev = torch.device("cuda")

def poisoned_empty(*shape, dtype, device):
    if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
        shape = tuple(shape[0])
    t = torch.empty(*shape, dtype=dtype, device=device)
    raw = t.view(-1).view(torch.uint8) if t.numel() else t
    if t.numel():
        n = raw.numel()
        pat = torch.tensor([0xC0, 0x7F], dtype=torch.uint8, device=device).repeat((n + 1) // 2)[:n]
        raw.copy_(pat)
    return t

KF._ALLOC = poisoned_empty
bad = 0
for precision in ("bf16", "fp32"):
    K.set_precision(precision)
    for (cin, cout, hw, n) in [(3, 64, 64, 4), (64, 64, 64, 4), (64, 128, 32, 4), (128, 128, 32, 4), (128, 256, 16, 4), (256, 256, 16, 4),
                               (256, 512, 8, 4), (512, 512, 8, 4), (512, 512, 4, 4), (512, 512, 4, 2), (64, 64, 224, 1), (3, 64, 37, 3), (24, 40, 9, 5)]:
        if precision == "fp32" and hw > 64:
            continue
        torch.manual_seed(0)
        m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU).to(dev)
        x = torch.randn(n, cin, hw, hw, device=dev, requires_grad=True)
        y = m(x)
        y.backward(torch.randn_like(y))
        torch.cuda.synchronize()
        res = {"y": y, "dx": x.grad}
        res.update({"d " + k: p.grad for k, p in m.named_parameters()})
        nan = [k for k, v in res.items() if v is not None and not bool(torch.isfinite(v).all())]
        bad += len(nan)
        print(f"{precision} {cin:4d} -> {cout:4d} @ {hw:3d} x{n}: " + ("clean" if not nan else "NaN in " + ", ".join(nan)), flush=True)
# whole model (adds the pooling kernels and the layer-to-layer buffer reuse)
from kanconv_b200.models.kan_vgg import vggkan
K.set_precision("bf16")
torch.manual_seed(0)
model = vggkan(3, 1000, arch="VGG16", classifier_type="Linear", expected_feature_shape=(7, 7), spline_order=3, grid_size=5).to(dev)
model.eval()
for n, hw in ((4, 64), (2, 64), (3, 96)):
    x = torch.randn(n, 3, hw, hw, device=dev)
    model.zero_grad(set_to_none=True)
    out = model(x)
    torch.nn.functional.cross_entropy(out, torch.randint(0, 1000, (n,), device=dev)).backward()
    torch.cuda.synchronize()
    nan = [k for k, p in model.named_parameters() if not bool(torch.isfinite(p.grad).all())]
    if not bool(torch.isfinite(out).all()): nan.insert(0, "logits")
    bad += len(nan)
    print(f"KAN-VGG16 bf16 @ {hw} x{n}: " + ("clean" if not nan else f"NaN in {len(nan)} results: " + ", ".join(nan[:8])), flush=True)
print("uninitialised reads found" if bad else "no uninitialised read found")
sys.exit(1 if bad else 0)
