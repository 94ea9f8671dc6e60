"""RayBundle / RaySamples / Frustums with the reference's attribute names and shapes
(NS/cameras/rays.py:33-295), backed by compact per-ray edge arrays for the CUDA operators."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import torch

from . import ops


@dataclass
class Frustums:
    """origins/directions [..., 3] (per ray [B,1,3] or per sample [B,S,3]); starts/ends/pixel_area [..., 1]."""

    origins: torch.Tensor
    directions: torch.Tensor
    starts: torch.Tensor
    ends: torch.Tensor
    pixel_area: Optional[torch.Tensor] = None
    offsets: Optional[torch.Tensor] = None
    _iv: Optional[ops.Intervals] = None  # compact euclidean edges when produced by our samplers

    @property
    def shape(self):
        return self.starts.shape[:-1]

    def intervals(self) -> ops.Intervals:
        if self._iv is None:
            s = self.starts[..., 0]
            e = self.ends[..., 0]
            if s.dim() == 1:
                s, e = s[:, None], e[:, None]
            self._iv = ops.Intervals(starts=s.reshape(-1, s.shape[-1]), ends=e.reshape(-1, e.shape[-1]))
        return self._iv

    def get_positions(self) -> torch.Tensor:
        """o + d*(start+end)/2 (NS/cameras/rays.py:49-58)."""
        iv = self.intervals()
        o, d = self.origins, self.directions
        shape = self.starts.shape[:-1]
        per_ray = o.dim() == len(shape) + 1 and (o.shape[-2] == 1 or len(shape) == 1)
        if per_ray and len(shape) >= 2:
            pos = ops.sample_positions(o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous(), iv)
        else:  # per-sample origins: every sample is its own single-sample ray
            o2 = o.expand(*shape, 3).reshape(-1, 3).contiguous()
            d2 = d.expand(*shape, 3).reshape(-1, 3).contiguous()
            flat = ops.Intervals(starts=self.starts.reshape(-1, 1), ends=self.ends.reshape(-1, 1))
            pos = ops.sample_positions(o2, d2, flat)
        pos = pos.view(*shape, 3)
        if self.offsets is not None:
            pos = pos + self.offsets
        return pos

    def get_start_positions(self) -> torch.Tensor:
        return self.origins + self.directions * self.starts

    def set_offsets(self, offsets):
        self.offsets = offsets


@dataclass
class RaySamples:
    frustums: Frustums
    camera_indices: Optional[torch.Tensor] = None
    _deltas: Optional[torch.Tensor] = None
    spacing_starts: Optional[torch.Tensor] = None
    spacing_ends: Optional[torch.Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    metadata: Optional[Dict[str, torch.Tensor]] = None
    times: Optional[torch.Tensor] = None
    _sdist: Optional[torch.Tensor] = None  # compact spacing edges [B,S+1]
    _nears: Optional[torch.Tensor] = None
    _fars: Optional[torch.Tensor] = None

    @property
    def shape(self):
        return self.frustums.shape

    @property
    def deltas(self) -> torch.Tensor:
        """ends - starts [B,S,1] (rays.py:291), evaluated on first use: no kernel on the hot path reads it."""
        if self._deltas is None:
            self._deltas = self.frustums.ends - self.frustums.starts
        return self._deltas

    def sdist(self) -> torch.Tensor:
        """[B,S+1] spacing edges = cat[spacing_starts, spacing_ends[-1]] (NS/model_components/losses.py:84-90)."""
        if self._sdist is None:
            self._sdist = torch.cat([self.spacing_starts[..., 0], self.spacing_ends[..., -1:, 0]], dim=-1).contiguous()
        return self._sdist

    def get_weights(self, densities: torch.Tensor) -> torch.Tensor:
        """alpha-compositing weights [B,S,1] (NS/cameras/rays.py:128-150)."""
        iv = self.frustums.intervals()
        w = ops.weights_from_density(densities.reshape(iv.B, iv.S), iv)
        return w.view(*densities.shape)


@dataclass
class RayBundle:
    origins: torch.Tensor
    directions: torch.Tensor
    pixel_area: Optional[torch.Tensor] = None
    camera_indices: Optional[torch.Tensor] = None
    nears: Optional[torch.Tensor] = None
    fars: Optional[torch.Tensor] = None
    metadata: Dict[str, torch.Tensor] = field(default_factory=dict)
    times: Optional[torch.Tensor] = None
    pose_corrected: bool = False  # the camera optimizer's correction is already applied (fused step prologue, data.py)

    def __len__(self) -> int:
        return self.origins.numel() // self.origins.shape[-1]

    def _map(self, fn) -> "RayBundle":
        f = lambda t: None if t is None else fn(t)
        return RayBundle(f(self.origins), f(self.directions), f(self.pixel_area), f(self.camera_indices), f(self.nears), f(self.fars),
                         {k: fn(v) for k, v in self.metadata.items()}, f(self.times), self.pose_corrected)

    def flatten(self) -> "RayBundle":
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def __getitem__(self, idx) -> "RayBundle":
        return self._map(lambda t: t[idx])

    def to(self, device) -> "RayBundle":
        return self._map(lambda t: t.to(device))

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        return self.flatten()[start_idx:end_idx]

    def get_ray_samples(self, sdist: torch.Tensor, ebins: torch.Tensor, spacing_to_euclidean_fn: Optional[Callable] = None) -> RaySamples:
        """Samples from compact edge arrays (our analogue of RayBundle.get_ray_samples, rays.py:251-295)."""
        iv = ops.Intervals(ebins=ebins)
        starts, ends = ebins[:, :-1, None], ebins[:, 1:, None]
        fr = Frustums(origins=self.origins[:, None, :], directions=self.directions[:, None, :], starts=starts, ends=ends,
                      pixel_area=None if self.pixel_area is None else self.pixel_area[:, None, :], _iv=iv)
        return RaySamples(
            frustums=fr,
            camera_indices=None if self.camera_indices is None else self.camera_indices[:, None, :],
            spacing_starts=sdist[:, :-1, None],
            spacing_ends=sdist[:, 1:, None],
            spacing_to_euclidean_fn=spacing_to_euclidean_fn,
            metadata={k: v[:, None, :] for k, v in self.metadata.items()},
            times=None if self.times is None else self.times[:, None, :],
            _sdist=sdist,
            _nears=self.nears,
            _fars=self.fars,
        )
