"""Same-seed state_dict fingerprints of the reference's two BASELINE models (dev container only; imports /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_model_golden.py

Stores, per parameter/buffer key: shape, sum and sum of absolute values (fp64) -> tests/golden/model_fingerprints.json, plus a
tiny forward fixture of the KAN-MobileNetV2 (FastKAN) model for the GPU parity test."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.dont_write_bytecode = True
sys.path.insert(0, os.environ.get("KAN_REFERENCE", "/root/reference"))
from models.kan_vgg import vggkan  # noqa: E402
from models.kan_mobilenetv2 import mobilenet_v2_kan  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "vgg16_kansmall_linear": lambda: vggkan(3, 10, arch="VGG16_kansmall", classifier_type="Linear"),
    "vgg16_small_kan_head": lambda: vggkan(3, 10, arch="VGG16_small", classifier_type="KAN", kan_classifier="KAN"),
    "mbv2_fastkan_kan_small_w025": lambda: mobilenet_v2_kan(num_classes=10, width_mult=0.25, arch="kan_small", kan_conv="FastKAN",
                                                           classifier_type="Linear"),
    "mbv2_kan_default_w025": lambda: mobilenet_v2_kan(num_classes=10, width_mult=0.25, kan_conv="KAN", classifier_type="Linear"),
}


def fingerprint(m):
    return {k: [list(v.shape), float(v.double().sum()), float(v.double().abs().sum())] for k, v in m.state_dict().items()}


out = {}
for name, ctor in CASES.items():
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ctor()
    out[name] = {"name": m.name, "n_params": sum(p.numel() for p in m.parameters()), "state": fingerprint(m)}
    print(name, m.name, out[name]["n_params"])
with open(os.path.join(HERE, "model_fingerprints.json"), "w") as f:
    json.dump(out, f)

# small forward/backward fixture: FastKAN-MobileNetV2 (kan_small, width 0.25) on 4x3x32x32, dropout off.  The model is badly
# conditioned at initialisation (train-mode BatchNorm over 16 values in the last blocks): the reference's own fp32 run deviates
# from its fp64 run by ~3e-3 (logits) / ~1e-1 (gradients), so that self-noise is stored next to the fp64 golden.
def run(dtype):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mobilenet_v2_kan(num_classes=10, width_mult=0.25, arch="kan_small", kan_conv="FastKAN", classifier_type="Linear", dropout=0.0)
    m = m.to(dtype).train()
    torch.manual_seed(1)
    x = torch.randn(4, 3, 32, 32)
    y = m(x.to(dtype))
    loss = y.square().mean()
    loss.backward()
    return x, y.detach(), float(loss.detach()), {k: p.grad for k, p in m.named_parameters() if p.grad is not None}


keys = ["features.0.spline_conv.0.weight", "features.3.conv.0.base_conv.0.weight", "classifier.fc.weight"]
x, y64, l64, g64 = run(torch.float64)
_, y32, l32, g32 = run(torch.float32)
np.savez_compressed(os.path.join(HERE, "mbv2_fastkan_forward.npz"), x=x.numpy(), y=y64.numpy(), loss=l64, y32=y32.numpy(), loss32=l32,
                    **{"grad/" + k: g64[k].numpy() for k in keys}, **{"grad32/" + k: g32[k].numpy() for k in keys})
print("mbv2 fixture: y", tuple(y64.shape), "loss", l64, "ref fp32 self-noise y", float((y32.double() - y64).abs().max() / y64.abs().max()))
