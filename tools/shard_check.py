"""Does a sample's result depend on the batch it is computed in?  KAN-VGG16 on 64x64 images (the shape of bench.py's DDP
self-check), BF16 path, one GPU: activations and activation gradients of every KAN convolution layer for the first two images
computed inside a batch of four vs as a batch of two; then the data-parallel property mean(shard gradients) == whole-batch
gradient, before and after a few optimizer steps."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
from kanconv_b200.models.kan_vgg import vggkan
dev = torch.device("cuda")
K.set_precision("bf16")
torch.manual_seed(0)
model = vggkan(3, 1000, arch="VGG16", classifier_type="Linear", expected_feature_shape=(7, 7), spline_order=3, grid_size=5).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
lossf = torch.nn.CrossEntropyLoss(reduction="sum")
g = torch.Generator().manual_seed(1234)
xt = torch.randn(8, 3, 224, 224, generator=g).to(dev); yt = torch.randint(0, 1000, (8,), generator=g).to(dev)
layers = [(n, m) for n, m in model.named_modules() if type(m).__name__ == "KANConv2DLayer"]
for steps in (0, 6):
    for _ in range(steps):
        opt.zero_grad(set_to_none=True); torch.nn.functional.cross_entropy(model(xt), yt).backward(); opt.step()
    gs = torch.Generator().manual_seed(4321)
    xa = torch.randn(4, 3, 64, 64, generator=gs).to(dev); ya = torch.randint(0, 1000, (4,), generator=gs).to(dev)
    model.eval()
    rec = {}
    def run(xs, ys, tag):
        acts, hooks = {}, []
        for n, m in layers:
            def fh(mod, inp, out, n=n):
                acts[n] = out.detach().double().clone()
                out.register_hook(lambda gr, n=n: acts.__setitem__(n + ":grad", gr.detach().double().clone()))
            hooks.append(m.register_forward_hook(fh))
        model.zero_grad(set_to_none=True)
        lossf(model(xs), ys).backward()            # sum reduction: per-sample gradients do not depend on the batch size
        for h in hooks: h.remove()
        rec[tag] = acts
        return [p.grad.detach().double().clone() for p in model.parameters()]
    whole = run(xa, ya, "whole")
    a = run(xa[:2], ya[:2], "shard"); b = run(xa[2:], ya[2:], "shard2")
    num = sum(float((u + v - w).square().sum()) for u, v, w in zip(a, b, whole)); den = sum(float(w.square().sum()) for w in whole)
    print(f"after {steps} training steps: sum(shard gradients) vs whole-batch gradient, relative L2 {(num / den) ** 0.5:.2e}")
    worst = sorted(((float((u + v - w).norm() / w.norm().clamp_min(1e-30)), float(w.norm()), n) for (n, _), u, v, w in zip(model.named_parameters(), a, b, whole)), reverse=True)[:6]
    for r, nrm, n in worst:
        print(f"   parameter {n:40s} relative L2 {r:.2e} (gradient norm {nrm:.2e})")
    for n, _ in layers:
        for k in (n, n + ":grad"):
            for lo, tag in ((0, "shard"), (2, "shard2")):
                u, v = rec["whole"][k][lo:lo + 2], rec[tag][k]
                d = float((u - v).abs().max() / v.abs().max().clamp_min(1e-30))
                if d > 0:
                    print(f"   {k:28s} samples {lo}-{lo + 1} shape {tuple(v.shape)}  max|batch-of-4 - batch-of-2| / max = {d:.2e}")
    model.train()

