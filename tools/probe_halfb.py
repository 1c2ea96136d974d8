import sys, os, ctypes, torch, torch.nn as nn
sys.path.insert(0, '/root/repo')
import kanconv_b200 as K
lib = K._lib.load()
for shape in [(64,256,256,56),(64,512,512,28),(32,128,128,112)]:
    n,cin,cout,hw = shape
    m = K.KANConv2DLayer(cin, cout, 3, padding=1, base_activation=nn.SiLU).cuda()
    x = torch.randn(n,cin,hw,hw,device='cuda', requires_grad=True)
    for flag in (0,1):
        lib.kc_debug_flag(flag)
        for mode in ("fwd","fwdbwd"):
            def run():
                y = m(x)
                if mode=="fwdbwd": y.backward(torch.ones_like(y))
            for _ in range(2): run()
            torch.cuda.synchronize()
            e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            print(shape, "halfB" if flag else "full ", mode, round(e0.elapsed_time(e1)/5,3), "ms")
lib.kc_debug_flag(0)
