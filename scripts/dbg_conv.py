import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.nn.functional as F
from odwscl_b200 import capi
def q(t, s=8): return (t * s).round() / s
for (B,H,W,Cin,Cout,dil) in [(1,8,16,32,32,1),(2,19,27,64,64,1),(1,38,50,64,128,1),(1,20,128,128,256,1),(2,76,128,512,512,2)]:
    g = torch.Generator().manual_seed(1)
    x = q(torch.randn(B,H,W,Cin,generator=g)).cuda(); w = q(torch.randn(Cout,Cin,3,3,generator=g)*0.5,16).cuda(); b = q(torch.randn(Cout,generator=g)).cuda()
    wk = w.permute(0,2,3,1).contiguous()
    ref = F.conv2d(x.permute(0,3,1,2).double(), w.double(), b.double(), padding=dil, dilation=dil).permute(0,2,3,1).float()
    for rep in range(4):
        got = capi.conv3x3_nhwc(x, wk, b, dilation=dil, flags=(capi.CONV_RELU if rep % 2 else 0))
        e = ref.clamp_min(0) if rep % 2 else ref
        print((B,H,W,Cin,Cout,dil), rep, "maxerr", float((got-e).abs().max()))
