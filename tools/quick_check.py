"""Quick GPU sanity of the fused step against the CPU oracle (fp32 exact mode, bf16 mode) - development aid."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vae-rs_b200")):
    sys.path.insert(0, p)
import torch
import models
from dataset import grid_batch, synthetic_tiles, grid_patch_pair
from oracle import ref_oracle as O
from svrs_native.trainer import FusedCondTrainer

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = models.Cond_SRVAE(2, 64)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
model.to(dev).train()
lr, hr = synthetic_tiles(1, 256, seed=3)
y, x = grid_batch(lr.to(dev), hr.to(dev), 64)
yo, xo = O.grid_batch(lr, hr, 64)
print("grid exact:", torch.equal(y.cpu(), yo), torch.equal(x.cpu(), xo))
pb = grid_patch_pair(hr.to(dev), 64, torch.bfloat16)
print("nhwc f32 exact:", torch.equal(pb.f32.permute(0, 3, 1, 2).cpu(), xo), "bf16 err", float((pb.op.float().permute(0, 3, 1, 2).cpu() - xo).abs().max()))
B = 4
y, x, yo, xo = y[:B], x[:B], yo[:B], xo[:B]
eng = model._engine()
g = torch.Generator().manual_seed(5)
eu, ez = torch.randn(B, eng.Wu, generator=g), torch.randn(B, eng.Wz, generator=g)
for mode in ("fp32", "bf16"):
    torch.manual_seed(0)
    m = models.Cond_SRVAE(2, 64)
    m.load_state_dict(sd)
    m.to(dev).train()
    if mode == "bf16":
        m.set_compute_dtype(torch.bfloat16)
    tr = FusedCondTrainer(m)
    sdo = {k: v.clone() for k, v in sd.items()}
    gam = {"gammax": torch.tensor(1.0), "gammay": torch.tensor(1.0)}
    opt = O.AdamState()
    for it in range(3):
        t = tr.step(x, y, eu.to(dev), ez.to(dev)).cpu()
        ref = O.cond_train_step(sdo, gam, opt, 2, 64, xo, yo, eu, ez)
        msg = []
        for i, k in enumerate(["mse_x", "kld_u", "mse_y", "kld_z", "loss"]):
            a, b = float(t[i]), float(ref[k])
            msg.append(f"{k} {a:.4f}/{b:.4f} ({abs(a-b)/abs(b):.1e})")
        print(mode, it, " ".join(msg))
    # parameters after 3 steps
    worst = 0
    for k, v in m.state_dict().items():
        if v.dtype.is_floating_point and "num_batches" not in k:
            d = float((v.cpu() - sdo[k]).abs().max())
            worst = max(worst, d)
    print(mode, "max |param - oracle| after 3 steps:", worst)
# graph replay + step_tiles
torch.manual_seed(0)
m = models.Cond_SRVAE(2, 64); m.load_state_dict(sd); m.to(dev).train(); m.set_compute_dtype(torch.bfloat16)
tr = FusedCondTrainer(m)
lr8, hr8 = synthetic_tiles(8, 256, seed=3)
lr8, hr8 = lr8.to(dev), hr8.to(dev)
for it in range(6):
    t = tr.step_tiles(hr8, lr8, patch_size=64, use_graph=True)
print("step_tiles graph loss:", t.cpu().tolist())
torch.cuda.synchronize()
print("OK")
