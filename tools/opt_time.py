"""How long do the parts of a KAN-VGG16 training step that are NOT this library take (torch's fused AdamW, loss, head)?
CUDA events around opt.step() and around the whole step, batch 64 @224."""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kanconv_b200 as K
from kanconv_b200.models import kan_vgg
dev = torch.device("cuda")
torch.manual_seed(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
arch, hw, classes, def_batch, feat, _ = bench.WORKLOADS["kan_vgg16_224"]
K.set_precision("bf16")
model = kan_vgg.vggkan(3, classes, arch=arch, classifier_type="Linear", expected_feature_shape=feat, spline_order=3, grid_size=5).to(dev)
model.train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True)
lossf = nn.CrossEntropyLoss()
x = torch.randn(def_batch, 3, hw, hw, device=dev); y = torch.randint(0, classes, (def_batch,), device=dev)
nparam = sum(p.numel() for p in model.parameters())
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tot = [0.0, 0.0, 0.0]
for it in range(8):
    ev[0].record()
    opt.zero_grad(set_to_none=True)
    loss = lossf(model(x), y)
    ev[1].record()
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 3:
        for k in range(3):
            tot[k] += ev[k].elapsed_time(ev[k + 1]) / 5
print(f"parameters {nparam / 1e6:.1f} M; forward + loss {tot[0]:.2f} ms, backward {tot[1]:.2f} ms, AdamW(fused) step {tot[2]:.2f} ms "
      f"({7 * 4 * nparam / tot[2] / 1e6:.0f} GB/s of the 7 x 4 B per parameter it moves)")
