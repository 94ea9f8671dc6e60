"""Renderers with the reference's call signatures (NS/model_components/renderers.py:58-447,
NS/model_components/shaders.py:56-77).  Each class evaluates through the fused per-ray render kernel; the model calls
`render_all` once instead of five separate passes."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops
from .rays import RaySamples


def _w2(weights):
    return weights[..., 0] if weights.dim() == 3 else weights


def render_all(weights, ray_samples: RaySamples, rgb=None, normals=None, pred_normals=None, eval_mode=False, want_median=True):
    """-> (rgb[B,3], accumulation[B,1], expected_depth[B,1], median_depth[B,1], median_idx[B,1], normals_img[B,3], pred_normals_img[B,3]);
    the normal maps are already NormalsShader-coded ((n+1)/2)."""
    iv = ray_samples.frustums.intervals()
    return ops.render(_w2(weights), iv, rgb, normals, pred_normals, eval_mode, want_median)


class RGBRenderer(nn.Module):
    def __init__(self, background_color="last_sample") -> None:
        super().__init__()
        if background_color != "last_sample":
            raise NotImplementedError("nvo_b200 RGBRenderer implements background_color='last_sample' (nerfacto default)")
        self.background_color = background_color

    def forward(self, rgb, weights, ray_samples: Optional[RaySamples] = None, ray_indices=None, num_rays=None, background_color=None):
        if ray_indices is not None:
            raise NotImplementedError("packed samples are not on the nerfacto path")
        w = _w2(weights)
        B, S = w.shape
        iv = ray_samples.frustums.intervals() if ray_samples is not None else ops.Intervals(ebins=torch.zeros((B, S + 1), device=w.device))
        return ops.render(w, iv, rgb, None, None, not self.training, False)[0]


class AccumulationRenderer(nn.Module):
    @classmethod
    def forward(cls, weights, ray_indices=None, num_rays=None):
        w = _w2(weights)
        B, S = w.shape
        iv = ops.Intervals(ebins=torch.zeros((B, S + 1), device=w.device))
        return ops.render(w, iv, None, None, None, False, False)[1]


class DepthRenderer(nn.Module):
    def __init__(self, method: str = "median") -> None:
        super().__init__()
        if method not in ("median", "expected"):
            raise NotImplementedError(f"Method {method} not implemented")
        self.method = method

    def forward(self, weights, ray_samples: RaySamples, ray_indices=None, num_rays=None):
        out = ops.render(_w2(weights), ray_samples.frustums.intervals(), None, None, None, False, self.method == "median")
        return out[3] if self.method == "median" else out[2]


class NormalsRenderer(nn.Module):
    """Returns safe_normalize(sum w n) (renderers.py:427-447)."""

    @classmethod
    def forward(cls, normals, weights, normalize: bool = True):
        if not normalize:
            raise NotImplementedError("normalize=False is not used on the nerfacto path")
        w = _w2(weights)
        B, S = w.shape
        iv = ops.Intervals(ebins=torch.zeros((B, S + 1), device=w.device))
        coded = ops.render(w, iv, None, normals, None, False, False)[5]
        return coded * 2.0 - 1.0


class NormalsShader(nn.Module):
    @classmethod
    def forward(cls, normals, weights=None):
        normals = (normals + 1) / 2
        return normals * weights if weights is not None else normals
