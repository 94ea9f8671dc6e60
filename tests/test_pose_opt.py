"""Camera-pose optimisation inside the mapping step (NS/models/nerfacto.py:171,249,288-291,379-380,
NS/cameras/camera_optimizers.py:108-167, group "camera_opt" of nerf_vo/mapping/nerfstudio.py:93-100): d loss / d pose_adjustment through
every sampling level's sample positions and the field's direction encoding, against the CPU oracle's autograd — through plain autograd
(model API) and through MappingTrainer's gradient sink — and the group's Adam under ExponentialDecayScheduler against torch."""
import math

import numpy as np
import pytest
import torch

import nerfacto_oracle as O
from test_full_size_parity import _record, _setup

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K, B = 8, 2048


@pytest.fixture(scope="module")
def nv():
    import nerf_vo_b200

    return nerf_vo_b200


def _oracle_pose_grad(ocfg, P, rays, targets, jit, pose0, mode):
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    pose = pose0.clone().requires_grad_(True)
    o, d = O.apply_pose_correction(rays["origins"], rays["directions"], rays["camera_indices"], pose, mode)
    r2 = dict(rays)
    r2["origins"], r2["directions"] = o, d
    _, oL, total = O.mapping_step(Pg, ocfg, r2, targets, jit)  # runs total.backward()
    reg = pose[:, :3].norm(dim=-1).mean() * 1e-2 + pose[:, 3:].norm(dim=-1).mean() * 1e-3  # camera_optimizers.py:149-155
    reg.backward()
    return float(total + reg), pose.grad.detach(), {k: v.grad for k, v in Pg.items()}


def _model(nv, mode):
    ocfg, P, m, rays, targets, jit, rb, batch = _setup(nv, 14, 12, K, "fp16", B)
    cfg = m.config
    cfg.camera_optimizer_mode = mode
    m2 = nv.ExtendedNerfactoModel(cfg, num_train_data=K)
    m2.load_state_dict(P, strict=False)
    m2 = m2.to(DEV).train()
    g = torch.Generator().manual_seed(7)
    pose0 = torch.randn(K, 6, generator=g) * torch.tensor([2e-2] * 3 + [1e-2] * 3)
    with torch.no_grad():
        m2.camera_optimizer.pose_adjustment.copy_(pose0.to(DEV))
    return ocfg, P, m2, rays, targets, jit, rb, batch, pose0


@pytest.mark.parametrize("mode", ["SO3xR3", "SE3"])
def test_pose_gradient_vs_oracle_autograd(nv, mode):
    ocfg, P, m, rays, targets, jit, rb, batch, pose0 = _model(nv, mode)
    ototal, og, _ = _oracle_pose_grad(ocfg, P, rays, targets, jit, pose0, mode)
    _, ld, _ = m.get_train_loss_dict(rb, batch, [j.to(DEV) for j in jit])
    assert "camera_opt_regularizer" in ld
    total = sum(ld.values())
    total.backward()
    torch.cuda.synchronize()
    g = m.camera_optimizer.pose_adjustment.grad.detach().cpu()
    rel_l2 = float((g - og).norm() / og.norm())
    rel_max = float((g - og).abs().max() / og.abs().max())
    cos = float((g.flatten() @ og.flatten()) / (g.norm() * og.norm()))
    _record(f"pose_grad_{mode}", {"rel_l2": rel_l2, "rel_max": rel_max, "cos": cos, "loss_rel": abs(float(total) - ototal) / abs(ototal)})
    assert abs(float(total) - ototal) < 2e-3 * abs(ototal)
    # fp16 features, fp16 saved feature derivatives (the gradient w.r.t. positions goes through them): production-precision tolerance
    assert cos > 0.995 and rel_l2 < 6e-2, (rel_l2, rel_max, cos)


def test_trainer_pose_sink_matches_autograd(nv):
    """MappingTrainer accumulates the ray gradients on side streams (ops.ray_grad_sink) and applies the pose backward itself: same gradient
    as the autograd path, then one Adam step of the camera group under the exponential schedule against torch."""
    from nerf_vo_b200.trainer import MappingTrainer

    mode = "SO3xR3"
    ocfg, P, m, rays, targets, jit, rb, batch, pose0 = _model(nv, mode)
    _, ld, _ = m.get_train_loss_dict(rb, batch, [j.to(DEV) for j in jit])
    sum(ld.values()).backward()
    g_auto = m.camera_optimizer.pose_adjustment.grad.detach().clone()
    m.zero_grad(set_to_none=True)
    tr = MappingTrainer(m, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=False, camera_opt_lr=1e-4, camera_opt_lr_final=1e-5, max_num_iterations=100)
    assert tr.cam_group is not None
    tr.set_inputs({k: v.to(DEV) for k, v in rays.items()}, {k: v.to(DEV) for k, v in targets.items()}, [j.to(DEV) for j in jit])
    tr.model.before_train_iteration(10 ** 6)  # anneal = 1, as in get_train_loss_dict above
    tr._forward_backward()
    torch.cuda.synchronize()
    _, off, n = tr.cam_group
    g_sink = tr.grad[off:off + K * 6].view(K, 6)
    assert float((g_sink - g_auto).abs().max()) <= 2e-3 * float(g_auto.abs().max()), float((g_sink - g_auto).abs().max() / g_auto.abs().max())
    # three steps of the group against torch.optim.Adam + LambdaLR with the reference's schedule function (schedulers.py:122-138)
    p_ref = torch.nn.Parameter(pose0.clone())
    opt = torch.optim.Adam([p_ref], lr=1e-4, eps=1e-15)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda s: math.exp(math.log(1e-4) * (1 - min(s / 100, 1)) + math.log(1e-5) * min(s / 100, 1)) / 1e-4)
    for _ in range(3):
        tr.grad[off:off + K * 6].copy_(g_auto.reshape(-1))
        tr._optimizer_camera()
        p_ref.grad = g_auto.cpu().clone()
        opt.step()
        sched.step()
    torch.cuda.synchronize()
    got = m.camera_optimizer.pose_adjustment.detach().cpu()
    moved = float((p_ref.detach() - pose0).abs().max())
    assert moved > 0 and float((got - p_ref.detach()).abs().max()) < 1e-4 * moved + 1e-9
    assert int(tr.cam_step) == 3


def test_trainer_with_pose_opt_graph_runs_and_moves_poses(nv):
    from nerf_vo_b200.trainer import MappingTrainer

    ocfg, P, m, rays, targets, jit, rb, batch, pose0 = _model(nv, "SO3xR3")
    tr = MappingTrainer(m, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=True)
    tr.set_inputs({k: v.to(DEV) for k, v in rays.items()}, {k: v.to(DEV) for k, v in targets.items()}, [j.to(DEV) for j in jit])
    tr.capture(warmup=2)
    assert torch.equal(m.camera_optimizer.pose_adjustment.detach().cpu(), pose0)  # capture() leaves the training state untouched
    losses = [float(tr.train_step()) for _ in range(8)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert int(tr.cam_step) == 8
    assert float((m.camera_optimizer.pose_adjustment.detach().cpu() - pose0).abs().max()) > 1e-5
