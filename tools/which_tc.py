import sys, ctypes
sys.path.insert(0, '/root/repo')
import torch, torch.nn as nn
import kanconv_b200 as K
from kanconv_b200 import functional as KF
from kanconv_b200.models import kan_mobilenetv2
orig = KF._use_tc
log = []
def patched(lib, desc, precision):
    r = orig(lib, desc, precision)
    log.append((desc.n, desc.cin, desc.cout, desc.h, desc.w, desc.kh, desc.stride_h, desc.nb, r, "" if r else lib.kc_last_error().decode()[:90]))
    return r
KF._use_tc = patched
m = kan_mobilenetv2.mobilenet_v2_kan(num_classes=1000, kan_conv="FastKAN", kan_classifier="FastKAN").cuda()
x = torch.randn(8, 3, 224, 224, device="cuda")
y = m(x)
for l in log:
    if not l[8]: print(l)
print(len(log), "conv calls;", sum(1 for l in log if not l[8]), "not on tensor cores")
