"""Where does the bf16 error of mu_z come from?  Fake-quant emulation of encoder_x on the CPU."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vae-rs_b200"), os.path.join(ROOT, "tests")]
import torch, torch.nn.functional as F
import fixtures as FX
from oracle import ref_oracle as O

q = lambda t: t.to(torch.bfloat16).float()
model, sd = FX.build("cond", 2, 64, seed=12)
r = O.PortableRng(22)
x = r.rand(16, 4, 64, 64)

def enc(x, mode):
    """mode flags: 'in' round the input image; 'w' round weights; 'act' round every stored activation;
    'prebn32' keep the conv output that feeds a BN in fp32; 'bnout32' keep BN output fp32"""
    h = q(x) if "in" in mode else x
    W = (lambda k: q(sd[k])) if "w" in mode else (lambda k: sd[k])
    A = q if "act" in mode else (lambda t: t)
    for i in range(3):
        p = f"encoder_x.{i}"
        h = A(F.conv2d(h, W(p + ".conv.weight"), sd[p + ".conv.bias"], padding=1))
        h = F.conv2d(h, W(p + ".downsample.weight"), sd[p + ".downsample.bias"], stride=2, padding=1)
        hs = h if "prebn32" in mode else A(h)           # stored tensor; statistics always from the fp32 result
        mu = h.mean(dim=(0, 2, 3), keepdim=True); var = h.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
        if i == 0 and mode == "in,w,act":
            print("   |mean|/std of pre-BN tensors, block", i, (mu.abs() / var.sqrt()).flatten()[:8])
        h = F.relu((hs - mu) / (var + 1e-5).sqrt() * sd[p + ".bn.weight"].view(1, -1, 1, 1) + sd[p + ".bn.bias"].view(1, -1, 1, 1))
        h = h if "bnout32" in mode else A(h)
    for i in (3, 4, 5, 6):
        h = F.conv2d(h, W(f"encoder_x.{i}.weight"), sd[f"encoder_x.{i}.bias"], padding=1)
        if i != 6:
            h = A(h)
    return h

ref = enc(x, "")
for mode in ("in", "w", "in,w", "in,w,act", "in,w,act,prebn32", "in,w,act,prebn32,bnout32", "w,act", "w,act,prebn32"):
    e = enc(x, mode)
    print(f"{mode:28s} rel-L2 {float((e - ref).norm() / ref.norm()):.3e}  rel-to-max {float((e - ref).abs().max() / ref.abs().max()):.3e}")
