"""Model mirrors (models/kan_vgg.py, models/kan_mobilenetv2.py): same seed -> same state_dict as the reference models
(fingerprints generated from the unmodified reference by tests/golden/make_model_golden.py), same model names."""
import json
import os

import pytest
import torch

from kanconv_b200.models import mobilenet_v2_kan, vggkan
from _util import GOLDEN

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


@pytest.mark.parametrize("name", sorted(CASES))
def test_same_seed_state_dict_matches_reference_model(name):
    with open(os.path.join(GOLDEN, "model_fingerprints.json")) as f:
        ref = json.load(f)[name]
    torch.manual_seed(0)
    m = CASES[name]()
    assert m.name == ref["name"]
    assert sum(p.numel() for p in m.parameters()) == ref["n_params"]
    sd = m.state_dict()
    assert list(sd) == list(ref["state"]), "state_dict keys / order differ from the reference"
    for k, (shape, s, a) in ref["state"].items():
        v = sd[k].double()
        assert list(v.shape) == shape, k
        assert abs(float(v.sum()) - s) <= 1e-9 * max(1.0, abs(a)), k
        assert abs(float(v.abs().sum()) - a) <= 1e-9 * max(1.0, abs(a)), k


def test_vgg11_cfg_is_defined():
    from kanconv_b200.models import cfgs
    assert cfgs["VGG11"] == [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512]
    with pytest.raises(ValueError):
        vggkan(3, 10, arch="VGG99")
