"""Same-seed state_dict fingerprints of the reference's two BASELINE models (dev container only; imports /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_model_golden.py [mbv2|vgg|mlp ...]     (no argument = everything)

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
    "vgg16_kansmall_hermite": lambda: vggkan(3, 10, arch="VGG16_kansmall", kan_conv="HermiteKAN", classifier_type="Linear"),
    "vgg16_kansmall_legendre": lambda: vggkan(3, 10, arch="VGG16_kansmall", kan_conv="LegendreKAN", classifier_type="Linear"),
    "vgg16_kansmall_jacobi": lambda: vggkan(3, 10, arch="VGG16_kansmall", kan_conv="JacobiKAN", classifier_type="Linear"),
    "mbv2_gegenbauer_kan_small_w025": lambda: mobilenet_v2_kan(num_classes=10, width_mult=0.25, arch="kan_small",
                                                              kan_conv="GegenbauerKAN", classifier_type="Linear"),
    "mbv2_fastkan_rdw_kan_small_w025": lambda: mobilenet_v2_kan(num_classes=10, width_mult=0.25, arch="kan_small", kan_conv="FastKAN",
                                                               classifier_type="Linear", replace_depthwise=True),
}


def fingerprint(m):
    return {k: [list(v.shape), float(v.double().sum()), float(v.double().abs().sum())] for k, v in m.state_dict().items()}


out = {}
for name, ctor in (CASES.items() if not sys.argv[1:] else []):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ctor()
    out[name] = {"name": m.name, "n_params": sum(p.numel() for p in m.parameters()), "state": fingerprint(m)}
    print(name, m.name, out[name]["n_params"])
if not sys.argv[1:]:
    with open(os.path.join(HERE, "model_fingerprints.json"), "w") as f:
        json.dump(out, f)

# small forward/backward fixture: FastKAN-MobileNetV2 (kan_small, width 0.25) on 4x3x32x32, dropout off.  The model is badly
# conditioned at initialisation (train-mode BatchNorm over 16 values in the last blocks): the reference's own fp32 run deviates
# from its fp64 run by ~3e-3 (logits) / ~1e-1 (gradients), so that self-noise is stored next to the fp64 golden.
def run(dtype, **extra):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mobilenet_v2_kan(num_classes=10, width_mult=0.25, arch="kan_small", kan_conv="FastKAN", classifier_type="Linear",
                             dropout=0.0, **extra)
    m = m.to(dtype).train()
    torch.manual_seed(1)
    x = torch.randn(4, 3, 32, 32)
    y = m(x.to(dtype))
    loss = y.square().mean()
    loss.backward()
    return x, y.detach(), float(loss.detach()), {k: p.grad for k, p in m.named_parameters() if p.grad is not None}


keys = ["features.0.spline_conv.0.weight", "features.3.conv.0.base_conv.0.weight", "classifier.fc.weight"]
if not sys.argv[1:] or "mbv2" in sys.argv[1:]:
    x, y64, l64, g64 = run(torch.float64)
    _, y32, l32, g32 = run(torch.float32)
    np.savez_compressed(os.path.join(HERE, "mbv2_fastkan_forward.npz"), x=x.numpy(), y=y64.numpy(), loss=l64, y32=y32.numpy(), loss32=l32,
                        **{"grad/" + k: g64[k].numpy() for k in keys}, **{"grad32/" + k: g32[k].numpy() for k in keys})
    print("mbv2 fixture: y", tuple(y64.shape), "loss", l64, "ref fp32 self-noise y", float((y32.double() - y64).abs().max() / y64.abs().max()))


# the same model with `replace_depthwise=True` (models/kan_mobilenetv2.py:112-124): the 7 depthwise stages become FastKAN
# convolutions with groups == channels (up to 48 single-channel groups per layer at this width)
rdw_keys = ["features.0.spline_conv.0.weight", "features.2.conv.1.spline_conv.5.weight", "features.2.conv.1.base_conv.0.weight",
            "features.4.conv.1.spline_conv.17.weight", "classifier.fc.weight"]
if not sys.argv[1:] or "mbv2_rdw" in sys.argv[1:]:
    x, y64, l64, g64 = run(torch.float64, replace_depthwise=True)
    _, y32, l32, g32 = run(torch.float32, replace_depthwise=True)
    rdw_keys = [k for k in rdw_keys if k in g64]
    np.savez_compressed(os.path.join(HERE, "mbv2_fastkan_rdw_forward.npz"), x=x.numpy(), y=y64.numpy(), loss=l64, y32=y32.numpy(),
                        loss32=l32, **{"grad/" + k: g64[k].numpy() for k in rdw_keys},
                        **{"grad32/" + k: g32[k].numpy() for k in rdw_keys})
    print("mbv2 rdw fixture: y", tuple(y64.shape), "loss", l64, "keys", rdw_keys,
          "ref fp32 self-noise y", float((y32.double() - y64).abs().max() / y64.abs().max()))


# ---- round 2: KAN-VGG (the benched model family) and the KAN MLP head, forward + backward fixtures ------------------------
# KAN-VGG16_kansmall (13 KAN convolutions, tail maps 2x2 at 32x32 input) and KAN-VGG11 (BASELINE config 3; the cfg is absent
# upstream, SURVEY 8(d) C3, so it is injected into the reference's cfgs dict at run time exactly as the survey did), head
# dropout off, CrossEntropy on fixed labels.  Stored: logits / loss in fp64 and fp32, (sum, L2 norm) of EVERY parameter gradient
# in fp64 and fp32, and a few gradients in full.
import models.kan_vgg as ref_vgg  # noqa: E402
from models.kans import MLP_KAN_FACTORY as REF_MLP  # noqa: E402

ref_vgg.cfgs["VGG11"] = [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512]


def run_vgg(arch, dtype, batch, size=32):
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = vggkan(3, 10, arch=arch, classifier_type="Linear", dropout_linear=0.0)
    m = m.to(dtype).train()
    torch.manual_seed(1)
    x = torch.randn(batch, 3, size, size)
    t = torch.arange(batch) % 10
    y = m(x.to(dtype))
    loss = nn.functional.cross_entropy(y, t)
    loss.backward()
    return x, t, y.detach(), float(loss.detach()), {k: p.grad for k, p in m.named_parameters() if p.grad is not None}


def vgg_fixture(arch, fname, batch, full_keys, size=32):
    x, t, y64, l64, g64 = run_vgg(arch, torch.float64, batch, size)
    _, _, y32, l32, g32 = run_vgg(arch, torch.float32, batch, size)
    summ = {k: [float(v.sum()), float(v.norm())] for k, v in g64.items()}
    summ32 = {k: [float(v.double().sum()), float(v.double().norm())] for k, v in g32.items()}
    # the reference's own fp32-vs-fp64 deviation of every gradient, max|a-b| / max|b|: the yardstick for an fp32 implementation
    noise = {k: float((g32[k].double() - v).abs().max() / v.abs().max().clamp_min(1e-300)) for k, v in g64.items()}
    np.savez_compressed(os.path.join(HERE, fname), x=x.numpy(), t=t.numpy(), y=y64.numpy(), loss=l64, y32=y32.numpy(), loss32=l32,
                        gradsum=np.frombuffer(json.dumps(summ).encode(), dtype=np.uint8),
                        gradsum32=np.frombuffer(json.dumps(summ32).encode(), dtype=np.uint8),
                        gradnoise32=np.frombuffer(json.dumps(noise).encode(), dtype=np.uint8),
                        **{"grad/" + k: g64[k].numpy() for k in full_keys}, **{"grad32/" + k: g32[k].numpy() for k in full_keys})
    print(fname, "y", tuple(y64.shape), "loss", l64, "ref fp32 self-noise y", float((y32.double() - y64).abs().max() / y64.abs().max()),
          "params", len(g64))


if not sys.argv[1:] or "vgg" in sys.argv[1:]:
    vgg_fixture("VGG16_kansmall", "vgg16_kansmall_forward.npz", 2,
                ["features.0.spline_conv.0.weight", "features.0.base_conv.0.weight", "features.0.prelus.0.weight",
                 "features.7.spline_conv.0.weight", "features.16.base_conv.0.weight", "features.16.prelus.0.weight",
                 "classifier.1.weight", "classifier.1.bias"])
    vgg_fixture("VGG11", "vgg11_forward.npz", 2,
                ["features.0.spline_conv.0.weight", "features.0.base_conv.0.weight", "features.11.prelus.0.weight",
                 "classifier.1.weight", "classifier.1.bias"])
    # the same KAN-VGG16_kansmall at 128x128 (tail maps 8x8 instead of 2x2): InstanceNorm over 4 values makes the 32x32 case
    # ill-conditioned (rounding is amplified ~700x from input to logits); this one is the model-level
    # yardstick of the BF16 tensor-core path
    vgg_fixture("VGG16_kansmall", "vgg16_kansmall_128_forward.npz", 2,
                ["features.0.spline_conv.0.weight", "features.16.base_conv.0.weight", "classifier.1.weight", "classifier.1.bias"],
                size=128)

# KAN MLP head (models/kans.py:300-327 through MLP_KAN_FACTORY['KAN']): [20, 16, 10], no dropout
if not sys.argv[1:] or "mlp" in sys.argv[1:]:
    def run_mlp(dtype):
        torch.manual_seed(0)
        m = REF_MLP["KAN"]([20, 16, 10], dropout=0.0).to(dtype).train()
        torch.manual_seed(1)
        x = torch.randn(8, 20) * 1.2
        torch.manual_seed(2)
        g = torch.randn(8, 10)
        xx = x.to(dtype).requires_grad_(True)
        y = m(xx)
        y.backward(g.to(dtype))
        return x, g, y.detach(), xx.grad.detach(), {k: p.grad for k, p in m.named_parameters()}, {k: v.detach() for k, v in m.state_dict().items()}

    x, g, y64, dx64, g64, sd = run_mlp(torch.float64)
    _, _, y32, dx32, g32, _ = run_mlp(torch.float32)
    np.savez_compressed(os.path.join(HERE, "kan_mlp_forward.npz"), x=x.numpy(), g=g.numpy(), y=y64.numpy(), dx=dx64.numpy(), y32=y32.numpy(),
                        dx32=dx32.numpy(), **{"sd/" + k: v.float().numpy() for k, v in sd.items()},
                        **{"grad/" + k: v.numpy() for k, v in g64.items()}, **{"grad32/" + k: v.numpy() for k, v in g32.items()})
    print("kan_mlp fixture: y", tuple(y64.shape), "keys", list(sd))
