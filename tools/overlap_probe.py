#!/usr/bin/env python
"""Standalone probe (not product code): do two phases of the mapping step overlap when they run side by side on one GPU?
Each phase (a list of launches) is captured into a CUDA graph alone and next to the other on a second stream; the replays are timed
with CUDA events.  Phases: the proposal sampling chain of the forward, the main table scatter, the fields group's Adam, the proposal
scatters.  usage: python tools/overlap_probe.py [--rays 4096]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_vo_b200 as nv
from nerf_vo_b200.synthetic import synthetic_jitters, synthetic_rays
from nerf_vo_b200.trainer import MappingTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.manual_seed(0)
model = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(camera_optimizer_mode="off"), num_train_data=192).to(dev)
tr = MappingTrainer(model, num_rays=a.rays)
rays, targets = synthetic_rays(a.rays, num_images=192, seed=1234)
jit = synthetic_jitters(a.rays, seed=99)
tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in jit])
tr.capture(warmup=3)
for _ in range(5):
    tr.train_step()
torch.cuda.synchronize()

ops = nv.ops
enc = model.field.mlp_base.encoder
x = model.field._cache["x"].detach().clone()
N = x.shape[0]
table = enc.hash_table.detach()
with torch.no_grad():
    model.proposal_sampler._steps_since_update = 0
    bundle = model.set_nears_and_fars(tr._bundle())
    jitters = [tr.inputs[f"jitter{k}"] for k in range(3)]
    _, _, rs_list = model.proposal_sampler(bundle, density_fns=model.density_fns, jitters=jitters)
    xs = [ops.contract_normalize(rs.frustums.get_positions().reshape(-1, 3).contiguous())[0] for rs in rs_list[:2]]
dy_main = torch.randn(ops.tmh_numel(N, enc.spec.out_dim), device=dev)
d_main = torch.zeros_like(table)
pspecs = [pn.encoding.spec for pn in model.proposal_networks]
dy_p = [torch.randn(ops.tmh_numel(xs[i].shape[0], pspecs[i].out_dim), device=dev) for i in range(2)]
d_p = [torch.zeros_like(pn.encoding.hash_table.detach()) for pn in model.proposal_networks]
n_par = tr.groups[0][2]
p2, m2, v2 = tr.flat[:n_par].clone(), torch.zeros(n_par, device=dev), torch.zeros(n_par, device=dev)
g2, cnt = tr.grad[:n_par].clone(), torch.zeros(1, dtype=torch.int32, device=dev)


def ph_prop():
    with torch.no_grad():
        model.proposal_sampler(bundle, density_fns=model.density_fns, jitters=jitters)


def ph_main_scatter():
    ops.grid_backward(x, dy_main, enc.spec, dtable=d_main, tmf=True)


def ph_adam():
    ops.adam_step(p2, g2, m2, v2, cnt, 1e-2, 0.9, 0.999, 1e-15)


def ph_prop_scatter(i):
    return lambda: ops.grid_backward(xs[i], dy_p[i], pspecs[i], dtable=d_p[i], tmf=True)


def ph_grid_fwd():
    ops.grid_forward_jac(x, table, enc.spec)


def seq(*fns):
    def f():
        for fn in fns:
            fn()
    return f


side = [torch.cuda.Stream() for _ in range(3)]
hi = torch.cuda.Stream(priority=-1)


def graph_time(branches, reps=30, prio=None):
    """branches: callables run concurrently (first on the capture stream, the others on side streams)."""
    for b in branches:
        b()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        used = []
        for k, b in enumerate(branches[1:]):
            st = hi if (prio is not None and prio == k + 1) else side[k]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                b()
            used.append(st)
        branches[0]()
        for st in used:
            cur.wait_stream(st)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


cases = {
    "prop chain (sampler forward)": [ph_prop],
    "main scatter": [ph_main_scatter],
    "adam fields": [ph_adam],
    "main scatter -> adam (one stream)": [seq(ph_main_scatter, ph_adam)],
    "main scatter || adam": [ph_main_scatter, ph_adam],
    "prop chain || (main scatter -> adam)": [ph_prop, seq(ph_main_scatter, ph_adam)],
    "prop chain || (main scatter -> adam) [tail high priority]": ([ph_prop, seq(ph_main_scatter, ph_adam)], 1),
    "prop chain || main scatter || adam": [ph_prop, ph_main_scatter, ph_adam],
    "prop chain || main scatter": [ph_prop, ph_main_scatter],
    "prop chain || adam": [ph_prop, ph_adam],
    "prop0 scatter": [ph_prop_scatter(0)],
    "prop1 scatter": [ph_prop_scatter(1)],
    "prop0 scatter || prop1 scatter": [ph_prop_scatter(0), ph_prop_scatter(1)],
    "prop0 scatter || adam": [ph_prop_scatter(0), ph_adam],
    "main scatter || prop0 scatter": [ph_main_scatter, ph_prop_scatter(0)],
    "main scatter || prop0 scatter || prop1 scatter": [ph_main_scatter, ph_prop_scatter(0), ph_prop_scatter(1)],
    "main scatter || prop0 scatter || prop1 scatter || adam": [ph_main_scatter, ph_prop_scatter(0), ph_prop_scatter(1), ph_adam],
    "grid fwd": [ph_grid_fwd],
    "grid fwd || adam": [ph_grid_fwd, ph_adam],
}
for name, c in cases.items():
    prio = None
    if isinstance(c, tuple):
        c, prio = c
    print(f"{name:70s} {graph_time(c, prio=prio):8.1f} us", flush=True)
