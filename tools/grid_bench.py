#!/usr/bin/env python
"""Isolated timings of the main hash-grid kernels (CUDA events, L2 flushed between launches): forward (TMH), forward + saved Jacobian,
re-gathering input gradient, Jacobian-based input gradient, scatter.  usage: python tools/grid_bench.py [--n 196608] [--log2 19]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_vo_b200 as nv

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=196608)
ap.add_argument("--log2", type=int, default=19)
ap.add_argument("--step-positions", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
L = 16
spec = nv.ops.GridSpec(L, a.log2, tuple(float(s) for s in nv.ops.torch_level_scalings(L, 16, 2048)))
x = torch.rand(a.n, 3, device=dev)
if a.step_positions:
    # the final-level sample positions of a real mapping step (48 PDF-resampled samples per ray: clustered, unlike uniform noise)
    from nerf_vo_b200.synthetic import synthetic_jitters, synthetic_rays
    from nerf_vo_b200.trainer import MappingTrainer
    B = a.n // 48
    model = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(), num_train_data=192).to(dev)
    with torch.no_grad():
        for nme, prm in model.named_parameters():
            if "hash_table" in nme:
                prm.normal_(0, 0.1)  # "trained-like" tables: densities are not ~1 everywhere, so the PDF samples cluster
    tr = MappingTrainer(model, num_rays=B, use_cuda_graph=False)
    rays, targets = synthetic_rays(B, num_images=192, seed=1234)
    tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in synthetic_jitters(B, seed=99)])
    for _ in range(3):
        tr.train_step()
    x = model.field._cache["x"].detach().clone()
    a.n = x.shape[0]
    del tr, model
table = (torch.rand(L << a.log2, 2, device=dev) * 2 - 1) * 1e-3
dtable = torch.zeros_like(table)
dy = torch.randn(nv.ops.tmh_numel(a.n, 2 * L), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
feat, jac = nv.ops.grid_forward_jac(x, table, spec)


def timed(name, fn, reps=20, alg_bytes=None):
    evs = []
    for i in range(3 + reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    us = sum(p.elapsed_time(q) for p, q in evs[3:]) / reps * 1e3
    extra = f"  {alg_bytes / us / 1e3:.0f} GB/s algorithmic" if alg_bytes else ""
    print(f"{name:44s} {us:8.1f} us{extra}", flush=True)


n = a.n
timed("grid_forward (tmh)", lambda: nv.ops.grid_forward(x, table, spec, "tmh"), alg_bytes=n * (12 + 16 * 64 + 64))
timed("grid_forward_jac (tmh + fp16 dy_dx)", lambda: nv.ops.grid_forward_jac(x, table, spec), alg_bytes=n * (12 + 16 * 64 + 64 + 192))
timed("grid_backward_input (re-gather)", lambda: nv.ops.grid_backward_input(x, table, dy, spec, tmf=True), alg_bytes=n * (12 + 16 * 64 + 128 + 12))
timed("grid_jac_dx", lambda: nv.ops.grid_jac_dx(jac, dy, spec, n, -1.0), alg_bytes=n * (192 + 128 + 12))
ref = None
for run in ("0", "8", "16"):
    os.environ["NVO_GRID_BWD_RUN"] = run
    dtable.zero_()
    nv.ops.grid_backward(x, dy, spec, dtable=dtable, tmf=True)
    if ref is None:
        ref = dtable.clone()
    else:
        print(f"  run={run}: max |d - d_ref| / max|d_ref| = {float((dtable - ref).abs().max() / ref.abs().max()):.2e}")
    timed(f"grid_backward (scatter) NVO_GRID_BWD_RUN={run}", lambda: nv.ops.grid_backward(x, dy, spec, dtable=dtable, tmf=True), alg_bytes=n * (12 + 128 + 16 * 64))
os.environ.pop("NVO_GRID_BWD_RUN")
# per-level cost of the scatter (one-level grids with the level's resolution): where the time goes between coarse (many samples per vertex,
# same-address reductions) and fine (every sample its own 8 vertices) levels
tot = 0.0
for l in range(L):
    sp = nv.ops.GridSpec(1, a.log2, (spec.scalings[l],))
    dyl = torch.randn(nv.ops.tmh_numel(n, 2), device=dev)
    dtl = torch.zeros((1 << a.log2, 2), device=dev)
    timed(f"  scatter level {l:2d} alone (scale {spec.scalings[l]:7.1f})", lambda: nv.ops.grid_backward(x, dyl, sp, dtable=dtl, tmf=True), reps=10)
