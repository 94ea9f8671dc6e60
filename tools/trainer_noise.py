#!/usr/bin/env python
"""How far do two runs of the SAME 16 training steps drift apart?  (atomic scatter order + Adam with eps=1e-15 amplify rounding noise.)
Compares the reference-style loop with itself and with MappingTrainer in eager / graph mode, early fields optimizer on / off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nerf_vo_b200 as nv
import test_trainer as T
from nerf_vo_b200.trainer import MappingTrainer

rays, targets, jit = T._inputs()


def loop():
    return T._reference_style_loop(nv, T._small_model(nv), rays, targets, jit)


def trainer(graph, early):
    os.environ["NVO_EARLY_FIELDS_OPT"] = early
    model = T._small_model(nv).to(T.DEV)
    tr = MappingTrainer(model, num_rays=T.B, lr=1e-2, eps=1e-15, use_cuda_graph=graph, proposal_update="reference")
    start = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr.capture(warmup=1)
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(start[k])
    for t in (tr.exp_avg, tr.exp_avg_sq):
        t.zero_()
    for c in tr.step_counts:
        c.zero_()
    tr.iteration, tr._ssu = 0, 0
    tr.set_inputs({k: v.to(T.DEV) for k, v in rays.items()}, {k: v.to(T.DEV) for k, v in targets.items()}, [j.to(T.DEV) for j in jit])
    losses = [float(tr.train_step()) for _ in range(T.STEPS)]
    return losses, None, {k: v.detach().clone() for k, v in model.state_dict().items()}


def report(name, a, b, start):
    worst = (0.0, "")
    for k, v in a[2].items():
        if not v.dtype.is_floating_point or v.ndim == 0:
            continue
        moved = (v - start[k].to(v.device)).abs().max()
        far = float(((b[2][k] - v).abs() > 0.05 * moved + 1e-6).float().mean())
        worst = max(worst, (far, k))
    dl = max(abs(x - y) / max(abs(y), 1e-9) for x, y in zip(a[0], b[0]))
    print(f"{name:34s} worst far-fraction {worst[0]:.4f} ({worst[1]}), max loss rel diff {dl:.2e}", flush=True)


start = {k: v.detach().clone() for k, v in T._small_model(nv).state_dict().items()}
L1 = loop(); L2 = loop()
report("loop vs loop", L1, L2, start)
for graph in (False, True):
    for early in ("0", "1"):
        for rep in range(2):
            report(f"trainer graph={graph} early={early} #{rep}", L1, trainer(graph, early), start)
