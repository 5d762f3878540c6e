"""CPU tests (-m "not gpu"), world_size 2 over gloo: the data-parallel loss-scaling rule of the fused step reproduces the
single-process gradient on the global batch (NLL = sum, KL = batch mean -> SUM all-reduce with KL pre-scaled by 1/R)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vae-rs_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _make(B):
    g = torch.Generator().manual_seed(0)
    d = dict(x=torch.rand(B, 4, 8, 8, generator=g), y=torch.rand(B, 4, 4, 4, generator=g),
             a=torch.randn(B, 4, 8, 8, generator=g), b=torch.randn(B, 4, 4, 4, generator=g),
             m=torch.randn(B, 6, 32, generator=g))
    w = torch.tensor([0.7, -0.3, 0.5, 0.2, 0.9, -0.4, 0.1, 0.6], dtype=torch.float64)
    return d, w


def _terms(d, w, lo, hi):
    from oracle import ref_oracle as O
    s = slice(lo, hi)
    rx = torch.sigmoid(d["a"][s].double() * w[0])
    ry = torch.sigmoid(d["b"][s].double() * w[1])
    m = d["m"][s].double()
    return O.cond_loss(rx, d["x"][s].double(), ry, d["y"][s].double(), m[:, 0] * w[2], m[:, 1] * w[3], m[:, 2] * w[4],
                       m[:, 3] * w[5], m[:, 4] * w[6], (m[:, 5] * w[7]).clamp(-7, 7),
                       torch.tensor(0.9, dtype=torch.float64), torch.tensor(1.2, dtype=torch.float64))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from svrs_native import parallel as P
    B = 6
    d, w = _make(B)
    w = w.clone().requires_grad_(True)
    lo, hi = P.shard_bounds(B, rank, world)
    assert P.sample_offset(B, rank, world) == lo
    terms = _terms(d, w, lo, hi)
    scales = P.upstream_grad_scales(world)
    loss = sum(t * s for t, s in zip(terms, scales))
    loss.backward()
    g = w.grad.clone()
    dist.all_reduce(g, op=dist.ReduceOp.SUM)                 # what the fused step does with the flat gradient buffer
    tsum = torch.stack([t.detach() for t in terms])
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save(dict(grad=g, terms=torch.tensor(P.global_terms(list(tsum), world))), out)
    dist.destroy_process_group()


def test_ddp_loss_scaling_rule_matches_global_batch(tmp_path):
    out = str(tmp_path / "ddp.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    d, w = _make(6)
    w = w.clone().requires_grad_(True)
    terms = _terms(d, w, 0, 6)
    sum(terms).backward()
    torch.testing.assert_close(res["grad"], w.grad, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(res["terms"], torch.stack([t.detach() for t in terms]), rtol=1e-12, atol=1e-12)
    # plain DDP averaging (mean of per-rank losses) would NOT reproduce it: documents why the rule exists
    w2 = w.detach().clone().requires_grad_(True)
    (0.5 * (sum(_terms(d, w2, 0, 3)) + sum(_terms(d, w2, 3, 6)))).backward()
    assert not torch.allclose(w2.grad, w.grad, rtol=1e-3)


def test_shard_bounds_cover_batch():
    from svrs_native import parallel as P
    for n in (1, 7, 128, 1024):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    assert P.upstream_grad_scales(8) == [1.0, 0.125, 1.0, 0.125]


def _agree_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models
    import dataset as D
    m = models.VAE(2, 32)
    m.world, m.rank = world, rank
    # a callback that wants to stop on ONE rank stops every rank (BaseVAE.fit agrees the decision with a MAX all-reduce)
    got = [m._agree(rank == 1, "cpu"), m._agree(False, "cpu")]
    # the rank-aware loader: this rank's share of every global batch of the SAME shuffled order and crop draws
    lr, hr = D.synthetic_tiles(12, 64, seed=3)
    plan = D.RandomCropLoader(D.TileDataset(lr, hr), 6, 32, device="cpu", shuffle=True, seed=9, rank=rank, world=world).plan()
    torch.save(dict(agree=got, plan=[(o, lo, n) for o, lo, n in plan], dist_info=D._dist_info()), f"{out}.{rank}")
    dist.destroy_process_group()


def test_fit_plumbing_under_world_size_two(tmp_path):
    """Host-side data-parallel plumbing of the public API on gloo: stop decisions agree across ranks, the loaders see the
    process group's (rank, world), and the two ranks' shards tile the single-process batches."""
    import dataset as D
    out = str(tmp_path / "plumb")
    port = 29500 + ((os.getpid() + 777) % 2000)
    mp.spawn(_agree_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert r0["agree"] == [True, False] and r1["agree"] == [True, False]
    assert r0["dist_info"] == (0, 2) and r1["dist_info"] == (1, 2)
    lr, hr = D.synthetic_tiles(12, 64, seed=3)
    single = D.RandomCropLoader(D.TileDataset(lr, hr), 6, 32, device="cpu", shuffle=True, seed=9).plan()
    assert len(single) == len(r0["plan"]) == len(r1["plan"]) == 2
    for (o, lo, n), (o0, lo0, n0), (o1, lo1, n1) in zip(single, r0["plan"], r1["plan"]):
        assert torch.equal(torch.cat([o0, o1]), o) and (lo0, lo1) == (0, n // 2) and n0 == n1 == n


def test_param_store_layout_puts_the_encoders_last():
    """ParamStore orders the flat buffers by backward completion (decoders | u_to_z, prior heads | encoders) so that each
    early all-reduce bucket of the data-parallel step is ONE contiguous range; state_dict order is untouched."""
    import models
    from svrs_native.engine import ParamStore
    m = models.Cond_SRVAE(2, 64)
    names = [n for n, _ in m.named_parameters()]
    st = ParamStore(m, late_prefixes=("encoder_y.", "encoder_x.", "y_to_z."))
    assert sorted(st.names) == sorted(names) and len(st.names) == len(names)
    late = [n.startswith(("encoder_y.", "encoder_x.", "y_to_z.")) for n in st.names]
    first_late = late.index(True)
    assert all(late[first_late:]) and not any(late[:first_late])
    assert st.offsets[first_late] == st.early_end and st.total % 4 == 0 and all(o % 4 == 0 for o in st.offsets)
    heads = [i for i, n in enumerate(st.names) if n.startswith(("u_to_z.", "mu_u_y_to_z.", "logvar_u_y_to_z."))]
    assert heads == list(range(heads[0], first_late)), "prior heads + u_to_z must sit right before the encoders"
    assert [n.split(".")[0] for n in st.names[:heads[0]]].count("decoder_x") > 0
    assert {n.split(".")[0] for n in st.names[:heads[0]]} == {"decoder_x", "decoder_y"}
    assert list(m.state_dict().keys())[0].startswith("encoder_y")        # the module's own order is unchanged
    plain = ParamStore(m)
    assert plain.names == names and plain.early_end == plain.total
