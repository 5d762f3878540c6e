"""bf16 throughput mode (the mode of every BENCH / SCALE number) against the CPU ORACLE at the BASELINE.json configs'
named shapes (-m gpu), replacing round 1's bf16-vs-own-fp32 self comparison:

  config 2  VAE cr=2 P=64, batch 256
  config 3  CondVAE cr=2 P=64 grid mode at the bench shape: 8 tiles -> grid-patch kernel -> 128 patches -> fused step
  config 4  CondVAE cr=16 on 256x256 crops (batch 2 of the named 128: same layers, same 128^2 / 256^2 maps)
  config 5  sample(): S posterior draws of one LR patch, bf16

Every case injects the same eps into both sides.  Two error figures are printed for each tensor:
  rel-to-max = max|err| / max|ref|   (the metric of the fp32 tests)
  rel-L2     = ||err||_2 / ||ref||_2
north star: "forward activations and ELBO terms within 1e-3 relative in bf16".  ELBO terms: met (tolerances below are
<= 2x the measured error and <= 1e-3 for the NLL terms and the loss).  Activations: a bf16 operand carries 8 mantissa
bits (ulp/2 = 2^-9 = 2e-3 relative per element per layer), so 1e-3 on max|err| is below the format's own rounding noise;
what is asserted instead is (a) rel-to-max / rel-L2 bounds at <= 2x the measured figures and (b) that this path's bf16
error is not larger than the error the REFERENCE ITSELF makes on the same inputs when run the way the reference would
run bf16 on this GPU (torch.autocast(bfloat16) over its cuDNN kernels) - test_bf16_error_vs_reference_autocast.
"""
import pytest
import torch

import fixtures as FX
from helpers import report
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAMES8 = ["x_hat", "y_hat", "mu_z", "logvar_z", "mu_u", "logvar_u", "mu_z_uy", "logvar_z_uy"]


IMAGES = ("x_hat", "y_hat", "sample")
FAILS = []


def _errs(name, got, ref, tol=None):
    """Prints both error figures; records a failure (asserted by _done() so that every tensor of a case is reported)."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert torch.isfinite(got).all(), name
    if tol is None:
        tol = ACT_TOL["image" if name.split()[-1] in IMAGES else "latent"]
    e = got - ref
    rmax = float(e.abs().max() / (ref.abs().max() + 1e-30))
    rl2 = float(e.norm() / (ref.norm() + 1e-30))
    print(f"[parity-bf16] {name}: rel-to-max {rmax:.3e} (tol {tol[0]:.1e})  rel-L2 {rl2:.3e} (tol {tol[1]:.1e})")
    if not (rmax <= tol[0] and rl2 <= tol[1]):
        FAILS.append((name, rmax, rl2))
    return rmax, rl2


def _done():
    bad = list(FAILS)
    FAILS.clear()
    assert not bad, bad


def _term(name, got, ref, tol):
    got, ref = float(got), float(ref)
    rel = abs(got - ref) / max(abs(ref), 1e-30)
    print(f"[parity-bf16] {name}: cuda {got:.6f} oracle {ref:.6f} rel {rel:.3e} (tol {tol:.1e})")
    if not (got == got and rel <= tol):
        FAILS.append((name, got, ref, rel))


def _gam2():
    return {"gammax": torch.tensor(1.0), "gammay": torch.tensor(1.0)}


# Tolerances, each <= 2x the worst figure the B200 run printed (profiles/parity_bf16_r02.txt):
#   ELBO terms: NLL terms and the loss 1e-4 (measured <= 1e-5), KL terms 1e-3 (measured <= 4.5e-4) - inside the north star's 1e-3;
#   gradient norm 2e-3 (measured <= 7e-4);
#   activations (rel-to-max, rel-L2): decoder images 4e-3 / 1e-3 (measured <= 1.9e-3 / 3.5e-4),
#                                     encoder / prior heads 3e-2 / 3e-2 (measured <= 1.7e-2 / 1.7e-2).
# The heads' ~1e-2 is the operand format, not the kernels: rounding ONLY the weights of encoder_x to bf16 in an fp32 CPU
# evaluation already moves mu_z by 5.8e-3 rel-L2, only the input image 4.0e-3, all stored tensors 1.07e-2
# (tools/dbg/bf16_err_probe.py, profiles/bf16_error_budget_r02.txt) - and the reference's own bf16 mode
# (autocast over cuDNN) is at 1.4e-2 on the same tensor (test_bf16_error_vs_reference_autocast).
ACT_TOL = {"image": (4e-3, 1e-3), "latent": (3e-2, 3e-2)}
TOL = {c: dict(nll=1e-4, kl=1e-3, gn=2e-3) for c in ("config2", "config3", "config4")}


def test_config2_vae_b256_bf16_vs_oracle():
    from svrs_native.trainer import FusedVaeTrainer
    tol = TOL["config2"]
    B, P, cr = 256, 64, 2
    model, sd = FX.build("vae", cr, P, seed=11, device=DEV, dtype=torch.bfloat16)
    r = O.PortableRng(21)
    x = r.rand(B, 4, P, P)
    eps = r.randn(B, model._engine().Wd)
    osd = {k: v.clone() for k, v in sd.items()}
    torch.set_num_threads(max(1, torch.get_num_threads()))
    terms_o, outs_o, _ = O.vae_train_step(osd, {"gamma": torch.tensor(1.0)}, O.AdamState(), cr, P, x, eps, return_grads=True)
    model.train()
    with torch.no_grad():
        outs = model(x.to(DEV), eps.to(DEV))
    for n, got, ref in zip(["x_hat", "mu", "logvar"], outs, outs_o):
        _errs(f"config2 fwd {n}", got, ref)
    model2, _ = FX.build("vae", cr, P, seed=11, device=DEV, dtype=torch.bfloat16)
    model2.train()
    tr = FusedVaeTrainer(model2)
    t = tr.step(x.to(DEV), eps.to(DEV)).cpu()
    _term("config2 mse", t[0], terms_o["mse"], tol["nll"])
    _term("config2 kld", t[1], terms_o["kld"], tol["kl"])
    _term("config2 loss", t[4], terms_o["loss"], tol["nll"])
    _term("config2 grad norm", tr.grad_norm(), terms_o["grad_norm"], tol["gn"])
    _done()


def test_config3_grid128_bf16_vs_oracle():
    """The bench workload itself: 8 synthetic 256x256 tiles -> TMA grid-patch gather + normalise (bit-exact) -> 128 patches
    -> fused bf16 step, against the oracle's grid_batch + cond_train_step on the same tiles and eps."""
    from dataset import grid_patch_pair, synthetic_tiles
    from svrs_native.trainer import FusedCondTrainer
    tol = TOL["config3"]
    P, cr, T = 64, 2, 8
    lr, hr = synthetic_tiles(T, 256, seed=100)
    yo, xo = O.grid_batch(lr, hr, P)
    B = xo.shape[0]
    assert B == 128
    model, sd = FX.build("cond", cr, P, seed=12, device=DEV, dtype=torch.bfloat16)
    eng = model._engine()
    r = O.PortableRng(22)
    eu, ez = r.randn(B, eng.Wu), r.randn(B, eng.Wz)
    osd = {k: v.clone() for k, v in sd.items()}
    terms_o, outs_o, _ = O.cond_train_step(osd, _gam2(), O.AdamState(), cr, P, xo, yo, eu, ez, return_grads=True)
    # patches from the device kernel (fp32 NHWC target + bf16 NHWC operand in one launch): bit-exact with the oracle
    xb = grid_patch_pair(hr.to(DEV), P, torch.bfloat16)
    yb = grid_patch_pair(lr.to(DEV), P // 2, torch.bfloat16)
    assert torch.equal(xb.f32.permute(0, 3, 1, 2).cpu(), xo) and torch.equal(yb.f32.permute(0, 3, 1, 2).cpu(), yo)
    assert torch.equal(xb.op.float().cpu(), xb.f32.to(torch.bfloat16).float().cpu())
    model.train()
    with torch.no_grad():
        outs = model(xo.to(DEV), yo.to(DEV), eu.to(DEV), ez.to(DEV))
    for n, got, ref in zip(NAMES8, outs, outs_o):
        _errs(f"config3 fwd {n}", got, ref)
    model2, _ = FX.build("cond", cr, P, seed=12, device=DEV, dtype=torch.bfloat16)
    model2.train()
    tr = FusedCondTrainer(model2)
    t = tr.step(xb, yb, eu.to(DEV), ez.to(DEV)).cpu()
    for i, k in enumerate(["mse_x", "kld_u", "mse_y", "kld_z", "loss"]):
        _term(f"config3 {k}", t[i], terms_o[k], tol["kl"] if k.startswith("kld") else tol["nll"])
    _term("config3 grad norm", tr.grad_norm(), terms_o["grad_norm"], tol["gn"])
    _done()


def test_config4_cr16_p256_bf16_vs_oracle():
    from svrs_native.trainer import FusedCondTrainer
    tol = TOL["config4"]
    P, cr, B = 256, 16, 2
    model, sd = FX.build("cond", cr, P, seed=13, device=DEV, dtype=torch.bfloat16)
    eng = model._engine()
    r = O.PortableRng(23)
    x = r.rand(B, 4, P, P)
    y = torch.nn.functional.avg_pool2d(x, 2)
    eu, ez = r.randn(B, eng.Wu), r.randn(B, eng.Wz)
    osd = {k: v.clone() for k, v in sd.items()}
    terms_o, outs_o, _ = O.cond_train_step(osd, _gam2(), O.AdamState(), cr, P, x, y, eu, ez, return_grads=True)
    model.train()
    with torch.no_grad():
        outs = model(x.to(DEV), y.to(DEV), eu.to(DEV), ez.to(DEV))
    for n, got, ref in zip(NAMES8, outs, outs_o):
        _errs(f"config4 fwd {n}", got, ref)
    model2, _ = FX.build("cond", cr, P, seed=13, device=DEV, dtype=torch.bfloat16)
    model2.train()
    tr = FusedCondTrainer(model2)
    t = tr.step(x.to(DEV), y.to(DEV), eu.to(DEV), ez.to(DEV)).cpu()
    for i, k in enumerate(["mse_x", "kld_u", "mse_y", "kld_z", "loss"]):
        _term(f"config4 {k}", t[i], terms_o[k], tol["kl"] if k.startswith("kld") else tol["nll"])
    _term("config4 grad norm", tr.grad_norm(), terms_o["grad_norm"], tol["gn"])
    _done()


def test_config5_sample_bf16_vs_oracle(golden_dir):
    fx = FX.load(golden_dir, "cond_cr2_p64_b2")
    model, sd = FX.build(fx, device=DEV, dtype=torch.bfloat16)
    model.eval()
    eng = model._engine()
    r = O.PortableRng(3)
    S = 32
    eu, es = r.randn(1, eng.Wu), r.randn(S, eng.Wz)
    y = FX.inputs(fx)[1][1:2]
    ref = O.cond_sample({k: v.clone() for k, v in sd.items()}, 2, 64, y, eu, es, training=False)
    with torch.no_grad():
        got = model.sample(y.to(DEV), samples=S, eps_u=eu.to(DEV), eps_s=es.to(DEV))
    _errs("config5 bf16 sample", got, ref)
    _done()


def test_bf16_error_vs_reference_autocast(golden_dir):
    """The reference's own way to run bf16 on this GPU is torch.autocast(bfloat16) over its cuDNN kernels.  On the golden
    fixture's inputs, this path's bf16 forward must be at least as close to the fp32 reference values as that."""
    fx = FX.load(golden_dir, "cond_cr2_p64_b2")
    model, sd = FX.build(fx, device=DEV, dtype=torch.bfloat16)
    x, y = FX.inputs(fx)
    eu, ez = fx["eps"]
    model.train()
    with torch.no_grad():
        outs = model(x.to(DEV), y.to(DEV), eu.to(DEV), ez.to(DEV))
    dsd = {k: v.to(DEV) for k, v in sd.items()}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        outs_ac = O.cond_forward(dsd, fx["cr"], fx["P"], x.to(DEV), y.to(DEV), eu.to(DEV), ez.to(DEV), training=True)
    worse = []
    for n, got, ac in zip(NAMES8, outs, outs_ac):
        ref = fx["outputs"][n].double()
        e_ours = float((got.double().cpu() - ref).norm() / ref.norm())
        e_ac = float((ac.double().cpu() - ref).norm() / ref.norm())
        m_ours = float((got.double().cpu() - ref).abs().max() / ref.abs().max())
        m_ac = float((ac.double().cpu() - ref).abs().max() / ref.abs().max())
        print(f"[parity-bf16] {n}: rel-L2 ours {e_ours:.3e} vs reference-under-autocast {e_ac:.3e};  rel-to-max ours {m_ours:.3e} "
              f"vs {m_ac:.3e}")
        if e_ours > 1.25 * e_ac + 1e-4:
            worse.append(n)
    assert not worse, f"bf16 error above the reference's own autocast error for {worse}"
