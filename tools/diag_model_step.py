"""Prints the error of every output / loss / gradient of one mapping step against the golden vectors (GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import nerf_vo_b200 as nv
from test_gpu_parity import _build_model
from nerf_vo_b200.fields import FieldHeadNames

g = dict(np.load(os.path.join(ROOT, "tests/golden/model_step_small.npz")))
m, rb, batch, jit = _build_model(nv, g)
m.proposal_sampler.set_anneal(float(g["anneal"]))
rec = {}
orig = m.field.forward
def fwd(rs, compute_normals=False):
    fo = orig(rs, compute_normals=compute_normals)
    rec.update({k.value: v.detach() for k, v in fo.items()})
    return fo
m.field.forward = fwd
outputs, loss_dict, _ = m.get_train_loss_dict(rb, batch, jit)
for i in range(3):
    w = outputs["weights_list"][i][..., 0].detach().cpu()
    print(f"level{i} weights max-abs err {float((w - torch.from_numpy(g[f'level{i}.weights'])).abs().max()):.3e}  sdist err {float((outputs['ray_samples_list'][i].sdist().cpu() - torch.from_numpy(g[f'level{i}.sdist'])).abs().max()):.3e}")
for k in ("density", "rgb", "normals", "pred_normals"):
    ref = torch.from_numpy(g[f"field.{k}"])
    got = rec[k].cpu()
    print(f"field.{k}: max-abs err {float((got - ref).abs().max()):.3e} (ref max {float(ref.abs().max()):.3e})")
acc = torch.from_numpy(g["out.accumulation"])[:, 0]
for k in ("rgb", "accumulation", "expected_depth", "normals", "pred_normals", "depth", "prop_depth_0", "prop_depth_1"):
    e = (outputs[k].detach().cpu() - torch.from_numpy(g[f"out.{k}"])).abs().reshape(len(acc), -1).max(dim=1)[0]
    print(f"out.{k}: max err {float(e.max()):.3e}; max err on rays with acc>0.5: {float(e[acc > 0.5].max()):.3e}; acc at worst ray {float(acc[e.argmax()]):.3e}")
for k, v in loss_dict.items():
    print(f"loss.{k}: got {float(v):.8e} ref {float(g[f'loss.{k}']):.8e}")
sum(loss_dict.values()).backward()
for name, p in m.named_parameters():
    ref = torch.from_numpy(g[f"grad.{name}"])
    got = p.grad.cpu() if p.grad is not None else torch.zeros_like(ref)
    print(f"grad {name}: max-abs err {float((got - ref).abs().max()):.3e} / max {float(ref.abs().max()):.3e}")
