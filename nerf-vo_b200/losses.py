"""Loss functions with the reference's names and argument meaning (NS/model_components/losses.py)."""
from __future__ import annotations

from enum import Enum

import torch

from . import ops
from .rays import RaySamples

EPS = 1.0e-7


class DepthLossType(Enum):
    DS_NERF = 1
    URF = 2
    SPARSENERF_RANKING = 3


def ray_samples_to_sdist(ray_samples: RaySamples) -> torch.Tensor:
    return ray_samples.sdist()


def _w2(w):
    return w[..., 0] if w.dim() == 3 else w


def interlevel_loss(weights_list, ray_samples_list) -> torch.Tensor:
    """losses.py:93-130."""
    c = ray_samples_to_sdist(ray_samples_list[-1]).detach()
    w = _w2(weights_list[-1]).detach()
    loss = 0.0
    for ray_samples, weights in zip(ray_samples_list[:-1], weights_list[:-1]):
        loss = loss + ops.interlevel_loss_op(w, c, _w2(weights), ray_samples_to_sdist(ray_samples))
    return loss


def distortion_loss(weights_list, ray_samples_list) -> torch.Tensor:
    """losses.py:147-153."""
    return ops.distortion_loss_op(_w2(weights_list[-1]), ray_samples_to_sdist(ray_samples_list[-1]))


def depth_loss(weights, ray_samples: RaySamples, termination_depth, predicted_depth, sigma, directions_norm, is_euclidean: bool,
               depth_loss_type: DepthLossType = DepthLossType.DS_NERF) -> torch.Tensor:
    """losses.py:288-324 (DS-NeRF)."""
    if depth_loss_type != DepthLossType.DS_NERF:
        raise NotImplementedError("Provided depth loss type not implemented.")
    if is_euclidean:
        directions_norm = torch.ones_like(termination_depth)
    sig = float(sigma) if not torch.is_tensor(sigma) else float(sigma.reshape(-1)[0])
    return ops.depth_loss_op(_w2(weights), ray_samples.frustums.intervals(), termination_depth, directions_norm, sig)


def monosdf_normal_loss(normal_pred, normal_gt) -> torch.Tensor:
    """losses.py:327-342."""
    return ops.normal_loss_op(normal_pred, normal_gt)


def rgb_mse_loss(gt_rgb, pred_rgb) -> torch.Tensor:
    """nn.MSELoss()(gt, pred) of NS/models/nerfacto.py:362."""
    return ops.mse_loss_op(pred_rgb, gt_rgb)


def orientation_loss(weights, normals, viewdirs) -> torch.Tensor:
    """losses.py:200-211 (multiplier 0 in NeRF-VO; kept for the loss_dict)."""
    n_dot_v = (normals * (-viewdirs)[..., None, :]).sum(dim=-1)
    return (_w2(weights) * torch.fmin(torch.zeros_like(n_dot_v), n_dot_v) ** 2).sum(dim=-1)


def pred_normal_loss(weights, normals, pred_normals) -> torch.Tensor:
    """losses.py:214-221 (multiplier 0 in NeRF-VO; kept for the loss_dict)."""
    return (_w2(weights) * (1.0 - torch.sum(normals * pred_normals, dim=-1))).sum(dim=-1)
