#!/usr/bin/env python
"""BASELINE.json configs 2 and 4 at their named shapes (parity-test cases, not bench lines): one fused bf16 step and one
fp32 forward of the same model/inputs; prints the ELBO terms of both so that the bf16-vs-fp32 agreement can be read off.
  config 2: VAE cr=2 P=64, batch 256        config 4: Cond_SRVAE cr=16 P=256, batch 4 (of 128; same layers/maps)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "simple-vae-rs_b200"))
import torch
import models

def run(name, make, inputs):
    torch.manual_seed(0)
    m = make().cuda()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    out = {}
    for dt in (torch.float32, torch.bfloat16):
        m.load_state_dict(sd)
        m.set_compute_dtype(dt)
        m.train()
        tr = m._fused_trainer(torch.optim.Adam(m.parameters(), lr=1e-4))
        terms = tr.step(*inputs, use_graph=False)
        torch.cuda.synchronize()
        out[dt] = [float(t) for t in terms]
        m._trainer = None
    a, b = out[torch.float32], out[torch.bfloat16]
    rel = max(abs(x - y) / max(abs(x), 1e-6) for x, y in zip(a, b))
    print(f"{name}: fp32 terms {['%.4f' % v for v in a]}  bf16 terms {['%.4f' % v for v in b]}  max rel diff {rel:.2e}  finite={all(map(lambda v: v == v, a + b))}")

g = torch.Generator().manual_seed(1)
x = torch.rand(256, 4, 64, 64, generator=g).cuda()
run("config 2 (VAE cr=2 P=64 B=256)", lambda: models.VAE(2, 64), (x,))
hr = torch.rand(4, 4, 256, 256, generator=g)
lr = torch.nn.functional.avg_pool2d(hr, 2)
run("config 4 (Cond_SRVAE cr=16 P=256 B=4)", lambda: models.Cond_SRVAE(16, 256), (hr.cuda(), lr.cuda()))
