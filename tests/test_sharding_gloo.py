"""N>1 host logic on CPU: two gloo ranks shard a global ray batch, all-reduce their flat gradients and recover the
single-process gradient (mean-loss semantics), and agree on the proposal-update schedule.  No kernels involved."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _toy_loss(w, rays, targets):
    # stands in for the mapping step: every term is a MEAN over the local rays, as in the reference's losses
    pred = torch.tanh(rays["origins"] @ w[:3] + rays["directions"] @ w[3:6])
    return ((pred - targets["rgb"][:, 0]) ** 2).mean() + 0.01 * (pred.abs()).mean()


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nerf_vo_b200 as nv

    sh = nv.sharding
    g = torch.Generator().manual_seed(0)
    B = 64
    rays = {"origins": torch.randn(B, 3, generator=g), "directions": torch.randn(B, 3, generator=g)}
    targets = {"rgb": torch.rand(B, 3, generator=g)}
    jit = [torch.rand(B, 1, generator=g) for _ in range(3)]
    w = torch.randn(6, generator=g, dtype=torch.float64).float().requires_grad_(True)
    # full-batch reference gradient (what one process would compute)
    full = torch.autograd.grad(_toy_loss(w, rays, targets), w)[0]
    r, t, j = sh.shard_batch(rays, targets, jit, rank, world)
    assert r["origins"].shape[0] == B // world and j[0].shape[0] == B // world
    a, b = sh.shard_range(B, rank, world)
    assert torch.equal(r["origins"], rays["origins"][a:b])
    flat = torch.autograd.grad(_toy_loss(w, r, t), w)[0].clone()
    scale = sh.allreduce_gradient_(flat)
    assert scale == 1.0 / world
    ok = torch.allclose(flat * scale, full, rtol=1e-5, atol=1e-7)
    # the `updated` predicate is a function of step counters only -> rank-invariant
    flags = torch.tensor([float(sh.proposal_update_due(s, k)) for s in (0, 9, 10, 2500, 5000, 8000) for k in (0, 1, 3, 5, 6)])
    gathered = [torch.zeros_like(flags) for _ in range(world)]
    dist.all_gather(gathered, flags)
    ok = ok and all(torch.equal(gathered[0], x) for x in gathered)
    torch.save({"ok": bool(ok), "flags": flags}, os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_matches_single_process(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        out = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert out["ok"], r
    # schedule known answers: always due for step < 10; steady state (>= warm-up) due only when more than 5 steps have passed
    import nerf_vo_b200 as nv

    f = nv.sharding.proposal_update_due
    assert f(3, 0) and not f(8000, 5) and f(8000, 6) and not f(2500, 2) and f(2500, 3)


def test_shard_range_errors():
    import nerf_vo_b200 as nv

    assert nv.sharding.shard_range(4096, 3, 8) == (1536, 2048)
    with pytest.raises(ValueError):
        nv.sharding.shard_range(10, 0, 3)
    with pytest.raises(ValueError):
        nv.sharding.shard_range(8, 2, 2)


# ---- evaluation frames: row sharding + gather (config 5) ---------------------------------------------------------------
def _row_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nerf_vo_b200 as nv

    sh = nv.sharding
    H, W = 37, 5  # 37 rows over 2 ranks: 19 + 18 (ragged)
    full_c = (torch.arange(H * W * 3) % 251).to(torch.uint8).view(H, W, 3)
    full_d = torch.arange(H * W, dtype=torch.float32).view(H, W)
    lo, hi = sh.row_shard(H, rank, world)
    c = sh.gather_rows(full_c[lo:hi].clone(), H, world)
    d = sh.gather_rows(full_d[lo:hi].clone(), H, world)
    torch.save({"ok": bool(torch.equal(c, full_c) and torch.equal(d, full_d)), "rows": (lo, hi)}, os.path.join(tmp, f"row{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_row_sharded_frame(tmp_path):
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_row_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rows = []
    for r in range(2):
        out = torch.load(os.path.join(str(tmp_path), f"row{r}.pt"))
        assert out["ok"], r
        rows.append(out["rows"])
    assert rows == [(0, 19), (19, 37)]


def test_row_shard_covers_every_row_once():
    import nerf_vo_b200 as nv

    for H in (1, 7, 360, 680, 681):
        for world in (1, 2, 3, 4, 8):
            spans = [nv.sharding.row_shard(H, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:])) and all(lo <= hi for lo, hi in spans)
    assert nv.sharding.row_shard(680, 3, 8) == (255, 340)
