"""Print the tile geometry the tensor-core kernels choose for the KAN-VGG16 layers (CPU only, no GPU needed)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn as nn
import kanconv_b200 as K
from kanconv_b200 import functional as KF
lib = K._lib.load()
lib.kc_tc_geometry.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]
shapes = [(32, 64, 128, 112), (32, 128, 128, 112), (64, 3, 64, 224), (64, 64, 64, 224), (64, 64, 128, 112), (64, 128, 128, 112), (64, 128, 256, 56), (64, 256, 256, 56),
          (64, 256, 512, 28), (64, 512, 512, 28), (64, 512, 512, 14), (16, 512, 512, 14), (16, 64, 64, 224)]
for n, cin, cout, hw in shapes:
    m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU)
    d = KF._make_desc(m._spec, n, cin, hw, hw, cout, cin * hw * hw, cout * hw * hw)
    row = []
    for which in (0, 1):
        out = (ctypes.c_longlong * 8)()
        rc = lib.kc_tc_geometry(ctypes.byref(d), which, out)
        row.append(dict(zip(["nsub", "ntile", "n_nt", "na", "tps", "bst", "mtiles", "smem"], list(out))) if rc == 0 else None)
    for which, r in zip(("fwd  ", "dgrad"), row):
        ctas = r["mtiles"] * r["n_nt"]
        print((n, cin, cout, hw), which, r, f"ctas={ctas} waves={ctas / 148:.2f}")
