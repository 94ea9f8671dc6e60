"""Fused gradient exchange + Adam (csrc/exchange.cu, nerf-vo_b200/peer.py).

CPU: the slice arithmetic (host mirror vs the library).  GPU: the kernel at world_size 1 against torch.optim.Adam, and —
when the box has two GPUs — two ranks exchanging through NVLink peer memory against the NCCL all-reduce + replicated Adam arm."""
import ctypes
import os
import socket
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slice_ranges_cover_the_flat_buffer_cpu():
    import nerf_vo_b200 as nv
    from nerf_vo_b200 import peer

    lib = nv._lib.load()
    for n in (4, 8, 1000, 18_432_516, 73_000_000):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = ctypes.c_int64(), ctypes.c_int64()
                length = lib.nvo_exchange_slice(n, r, world, ctypes.addressof(lo), ctypes.addressof(hi))
                assert (lo.value, hi.value) == peer.slice_range(n, r, world)
                assert length == hi.value - lo.value and lo.value % 4 == 0
                assert lo.value == prev or lo.value == hi.value
                prev = max(prev, hi.value)
            assert prev == (n + 3) // 4 * 4


@pytest.mark.gpu
def test_exchange_world1_matches_torch_adam():
    import nerf_vo_b200 as nv
    from nerf_vo_b200.peer import PeerBuffers

    dev = torch.device("cuda", 0)
    n = 1 << 20
    bufs = PeerBuffers(n, dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    p0 = torch.randn(n, generator=g)
    bufs.params.copy_(p0)
    ref = p0.clone().to(dev).requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-2, eps=1e-15)
    step = torch.zeros(1, dtype=torch.int32, device=dev)
    for it in range(4):
        grad = torch.randn(n, generator=g).to(dev) * 10.0 ** (-it)
        bufs.grads.copy_(grad)
        bufs.adam_exchange_step(step, 1e-2, 0.9, 0.999, 1e-15)
        ref.grad = grad.clone()
        opt.step()
    torch.cuda.synchronize()
    assert int(step) == 4 and bufs.error_word() == 0
    # fp32 Adam, same formula; torch's foreach kernels may contract differently: tolerance 2e-6 absolute on O(1) parameters
    assert float((bufs.params - ref.detach()).abs().max()) < 2e-6


def _two_rank_worker(rank, world, port, tmp):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import nerf_vo_b200 as nv
    from nerf_vo_b200 import ops
    from nerf_vo_b200.peer import PeerBuffers

    n = 4 * 1_000_003  # odd number of float4s: ragged last slice
    bufs = PeerBuffers(n, dev)
    g = torch.Generator(device="cpu").manual_seed(7)
    p0 = torch.randn(n, generator=g)
    bufs.params.copy_(p0)
    # library arm: NCCL sum-all-reduce + replicated nvo_adam_step
    flat, m, v = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step_a = torch.zeros(1, dtype=torch.int32, device=dev)
    step_b = torch.zeros(1, dtype=torch.int32, device=dev)
    worst, hist = 0.0, []
    for it in range(5):
        gr = torch.Generator(device="cpu").manual_seed(100 * it + rank)
        grad = torch.randn(n, generator=gr).to(dev)
        bufs.grads.copy_(grad)
        bufs.adam_exchange_step(step_a, 1e-2, 0.9, 0.999, 1e-15)
        red = grad.clone()
        dist.all_reduce(red)
        ops.adam_step(flat, red, m, v, step_b, 1e-2, 0.9, 0.999, 1e-15, 1.0 / world)
        torch.cuda.synchronize()
        worst = max(worst, float((bufs.params - flat).abs().max()))
        hist.append(float((bufs.params - flat).abs().max()))
    # the gradient sum of two ranks is order-independent (step 1 is bit-identical on B200); later steps differ by 1 ulp of an O(4)
    # parameter because the two kernels contract mul+add into FMA differently: tolerance 1e-6 absolute
    ok = worst < 1e-6 and hist[0] == 0.0
    ok = ok and bufs.error_word() == 0
    # ---- two parameter groups: group A alone (its counter runs ahead), then both groups in one launch ----------------------------
    n_a = 4 * 700_001
    b2 = PeerBuffers(n, dev)
    ga, gb = b2.add_group(0, n_a), b2.add_group(n_a, n - n_a)
    b2.params.copy_(p0)
    f2, m2, v2 = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    ca, cb = (torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(2))
    ra, rb = (torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(2))
    worst2 = 0.0
    for it, both in enumerate([True, False, True, True]):
        gr = torch.Generator(device="cpu").manual_seed(900 + 10 * it + rank)
        grad = torch.randn(n, generator=gr).to(dev)
        b2.grads.copy_(grad)
        if both:
            b2.adam_exchange_groups2(ga, ca, gb, cb, 1e-2, 0.9, 0.999, 1e-15)
        else:
            b2.adam_exchange_group(ga, ca, 1e-2, 0.9, 0.999, 1e-15)
        red = grad.clone()
        dist.all_reduce(red)
        ops.adam_step(f2[:n_a], red[:n_a], m2[:n_a], v2[:n_a], ra, 1e-2, 0.9, 0.999, 1e-15, 1.0 / world)
        if both:
            ops.adam_step(f2[n_a:], red[n_a:], m2[n_a:], v2[n_a:], rb, 1e-2, 0.9, 0.999, 1e-15, 1.0 / world)
        torch.cuda.synchronize()
        worst2 = max(worst2, float((b2.params - f2).abs().max()))
    ok = ok and worst2 < 1e-6 and b2.error_word() == 0 and (int(ca), int(cb)) == (4, 3) == (int(ra), int(rb))
    lo, hi = bufs.slice
    own = float((bufs.params[lo:hi] - flat[lo:hi]).abs().max())
    torch.save({"ok": bool(ok), "worst": worst, "worst_groups": worst2, "group_steps": (int(ca), int(cb)), "per_step": hist, "own_slice_err": own, "error_word": bufs.error_word(), "steps": (int(step_a), int(step_b))},
               os.path.join(tmp, f"r{rank}.pt"))
    dist.barrier()
    b2.close()
    bufs.close()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_exchange_two_ranks_matches_nccl_arm(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_two_rank_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        out = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert out["ok"], (r, out)


def _two_rank_trainer_worker(rank, world, port, tmp):
    """16 steps of the mapping trainer on two ranks (fused peer-memory arm), undeferred and with defer_fields_update: replicas stay bit-identical,
    the deferred run reproduces the undeferred loss trajectory and, after flush(), its parameter movement."""
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nerfacto_oracle as O
    import nerf_vo_b200 as nv
    from nerf_vo_b200.trainer import MappingTrainer

    K, B, STEPS = 8, 256, 16
    rays, targets = O.synthetic_rays(B, num_images=K, seed=3 + rank)
    jit = O.synthetic_jitters(B)
    res = {}
    for defer in (False, True):
        torch.manual_seed(0)
        cfg = nv.NerfactoModelConfig(log2_hashmap_size=14)
        for a in cfg.proposal_net_args_list:
            a["log2_hashmap_size"] = 12
        model = nv.ExtendedNerfactoModel(cfg, num_train_data=K).to(dev)
        tr = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=True, exchange="fused", defer_fields_update=defer)
        init = tr.flat.detach().clone()
        tr.capture(warmup=2)
        tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in jit])
        losses = [float(tr.train_step()) for _ in range(STEPS)]
        tr.flush()
        torch.cuda.synchronize()
        dist.barrier()
        flat = tr.flat.detach().clone()
        other = flat.clone()
        dist.broadcast(other, 0)
        res[defer] = {"losses": losses, "move": (flat - init).double().cpu(), "replicas_equal": bool(torch.equal(other, flat)), "steps": [int(c) for c in tr.step_counts],
                      "error_word": tr.peer.error_word()}
        dist.barrier()
    d0, d1 = res[False]["move"], res[True]["move"]
    out = {"cos": float((d0 * d1).sum() / (d0.norm() * d1.norm())), "norm_rel": abs(float(d0.norm()) - float(d1.norm())) / float(d0.norm()),
           "losses": (res[False]["losses"], res[True]["losses"]), "replicas_equal": (res[False]["replicas_equal"], res[True]["replicas_equal"]),
           "steps": (res[False]["steps"], res[True]["steps"]), "error_words": (res[False]["error_word"], res[True]["error_word"])}
    torch.save(out, os.path.join(tmp, f"t{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_trainer_deferred_fields_update(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_two_rank_trainer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        out = torch.load(os.path.join(str(tmp_path), f"t{r}.pt"))
        assert out["replicas_equal"] == (True, True), out["replicas_equal"]
        assert out["error_words"] == (0, 0)
        assert out["steps"][0] == out["steps"][1] == [16, 16], out["steps"]
        assert out["losses"][1] == pytest.approx(out["losses"][0], rel=2e-3), out["losses"]
        assert out["cos"] > 0.995 and out["norm_rel"] < 2e-2, (out["cos"], out["norm_rel"])
