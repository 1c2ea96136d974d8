"""Where does the KAN-VGG model-level deviation come from?  Per-parameter gradient error of the FP32 path vs the fp64 oracle
(next to the reference's own fp32-vs-fp64 noise from the fixture), and the layer-by-layer growth of the BF16 deviation."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kanconv_b200 as K  # noqa: E402
from kanconv_b200.models import vggkan  # noqa: E402
from oracle import kan_oracle as O  # noqa: E402
from _util import rel_err  # noqa: E402

arch, fixture = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("VGG16_kansmall", "vgg16_kansmall_forward")
z = np.load(os.path.join(ROOT, "tests", "golden", fixture + ".npz"))
gs, gs32 = json.loads(bytes(z["gradsum"]).decode()), json.loads(bytes(z["gradsum32"]).decode())
x, t = torch.from_numpy(z["x"]), torch.from_numpy(z["t"])
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False

torch.manual_seed(0)
ora = O.OracleVGG(3, 10, arch=arch, dropout_linear=0.0).double().train()
acts_o = []
hooks = [m.register_forward_hook(lambda m, i, o: acts_o.append(o.detach())) for m in ora.features]
F.cross_entropy(ora(x.double()), t).backward()
for h in hooks:
    h.remove()
go = {k: p.grad for k, p in ora.named_parameters()}

for prec in ("fp32", "auto"):
    torch.manual_seed(0)
    m = vggkan(3, 10, arch=arch, classifier_type="Linear", dropout_linear=0.0).cuda().train()
    K.set_precision(prec)
    acts = []
    hooks = [l.register_forward_hook(lambda m_, i, o: acts.append(o.detach())) for l in m.features]
    y = m(x.cuda())
    F.cross_entropy(y, t.cuda()).backward()
    torch.cuda.synchronize()
    K.set_precision("auto")
    print(f"== {arch} precision {prec}: logits err {rel_err(y, torch.from_numpy(z['y'])):.2e}")
    print("   activation error per feature layer:", " ".join(f"{rel_err(a, b):.1e}" for a, b in zip(acts, acts_o)))
    for k, p in m.named_parameters():
        nrm = gs[k][1]
        print(f"   {k:38s} full {rel_err(p.grad, go[k]):.2e}  norm {abs(float(p.grad.double().norm()) - nrm) / max(nrm, 1e-30):.2e}"
              f"  (reference fp32 norm noise {abs(gs32[k][1] - nrm) / max(nrm, 1e-30):.2e})  |g| {nrm:.2e}")
