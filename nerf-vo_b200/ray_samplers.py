"""Samplers with the reference's API (NS/model_components/ray_samplers.py:30-128,225-372,523-618)."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .rays import RayBundle, RaySamples


class Sampler(nn.Module):
    def __init__(self, num_samples: Optional[int] = None) -> None:
        super().__init__()
        self.num_samples = num_samples

    def forward(self, *args, **kwargs):
        return self.generate_ray_samples(*args, **kwargs)


def _piecewise_spacing_to_euclidean(nears, fars) -> Callable:
    """Host-visible closure equivalent to ray_samplers.py:112-117 (kept for API parity; kernels evaluate it on the device).
    Lazy: s_near / s_far are only computed if somebody calls it."""
    f = lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x))
    finv = lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))

    def fn(x):
        s_near, s_far = f(nears), f(fars)
        return finv(x * s_far + (1 - x) * s_near)

    return fn


class UniformLinDispPiecewiseSampler(Sampler):
    """ray_samplers.py:225-248: first half uniform, second half linear in disparity; single jitter per ray in training."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False) -> None:
        super().__init__(num_samples=num_samples)
        if train_stratified and not single_jitter:
            raise NotImplementedError("nvo_b200 implements the single-jitter stratified sampler nerfacto uses (use_single_jitter=True)")
        self.train_stratified, self.single_jitter = train_stratified, single_jitter

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, num_samples: Optional[int] = None, jitter: Optional[torch.Tensor] = None) -> RaySamples:
        """`jitter` [B,1] overrides the internally drawn torch.rand (lets tests share randomness with the oracle)."""
        assert ray_bundle is not None and ray_bundle.nears is not None and ray_bundle.fars is not None
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        B = ray_bundle.origins.shape[0]
        if self.train_stratified and self.training:
            if jitter is None:
                jitter = torch.rand((B, 1), dtype=torch.float32, device=ray_bundle.origins.device)
        else:
            jitter = None
        sdist, ebins = ops.sample_uniform(num_samples, ray_bundle.nears, ray_bundle.fars, jitter)
        return ray_bundle.get_ray_samples(sdist, ebins, _piecewise_spacing_to_euclidean(ray_bundle.nears, ray_bundle.fars))


class PDFSampler(Sampler):
    """ray_samplers.py:251-372 with include_original=False."""

    def __init__(self, num_samples: Optional[int] = None, train_stratified: bool = True, single_jitter: bool = False, include_original: bool = True,
                 histogram_padding: float = 0.01) -> None:
        super().__init__(num_samples=num_samples)
        if include_original:
            raise NotImplementedError("include_original=True is not on the nerfacto path (ProposalNetworkSampler passes False)")
        if train_stratified and not single_jitter:
            raise NotImplementedError("nvo_b200 implements the single-jitter PDF sampler nerfacto uses")
        self.train_stratified, self.single_jitter = train_stratified, single_jitter
        self.include_original, self.histogram_padding = include_original, histogram_padding

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, ray_samples: Optional[RaySamples] = None, weights: Optional[torch.Tensor] = None,
                             num_samples: Optional[int] = None, eps: float = 1e-5, jitter: Optional[torch.Tensor] = None, anneal: float = 1.0,
                             return_inds: bool = False):
        if ray_samples is None or ray_bundle is None:
            raise ValueError("ray_samples and ray_bundle must be provided")
        assert weights is not None, "weights must be provided"
        num_samples = num_samples or self.num_samples
        assert num_samples is not None
        B = ray_bundle.origins.shape[0]
        if self.train_stratified and self.training:
            if jitter is None:
                jitter = torch.rand((B, 1), dtype=torch.float32, device=weights.device)
        else:
            jitter = None
        w = weights[..., 0] if weights.dim() == 3 else weights
        res = ops.pdf_resample(w, ray_samples.sdist(), num_samples, ray_bundle.nears, ray_bundle.fars, jitter, anneal, self.histogram_padding, return_inds)
        out = ray_bundle.get_ray_samples(res[0], res[1], ray_samples.spacing_to_euclidean_fn)
        return (out, res[2]) if return_inds else out


class ProposalNetworkSampler(Sampler):
    """ray_samplers.py:523-618."""

    def __init__(self, num_proposal_samples_per_ray: Tuple[int, ...] = (64,), num_nerf_samples_per_ray: int = 32, num_proposal_network_iterations: int = 2,
                 single_jitter: bool = False, update_sched: Callable = lambda x: 1, initial_sampler: Optional[Sampler] = None,
                 pdf_sampler: Optional[PDFSampler] = None) -> None:
        super().__init__()
        self.num_proposal_samples_per_ray = num_proposal_samples_per_ray
        self.num_nerf_samples_per_ray = num_nerf_samples_per_ray
        self.num_proposal_network_iterations = num_proposal_network_iterations
        self.update_sched = update_sched
        if self.num_proposal_network_iterations < 1:
            raise ValueError("num_proposal_network_iterations must be >= 1")
        self.initial_sampler = initial_sampler if initial_sampler is not None else UniformLinDispPiecewiseSampler(single_jitter=single_jitter)
        self.pdf_sampler = pdf_sampler if pdf_sampler is not None else PDFSampler(include_original=False, single_jitter=single_jitter)
        self._anneal = 1.0
        self._anneal_dev: Optional[torch.Tensor] = None
        self._steps_since_update = 0
        self._step = 0

    def set_anneal(self, anneal: float) -> None:
        changed = float(anneal) != float(self._anneal)
        self._anneal = anneal
        if self._anneal_dev is not None and changed:  # past the anneal window the value stays 1.0: no launch
            self._anneal_dev.fill_(float(anneal))

    def use_device_anneal(self, device) -> torch.Tensor:
        """Keeps the anneal value in a device scalar the resampling kernel reads: a CUDA graph captured around the sampler then follows
        set_anneal() on every replay (as a by-value kernel argument it would stay frozen at its capture-time value)."""
        if self._anneal_dev is None or self._anneal_dev.device != torch.device(device):
            self._anneal_dev = torch.full((1,), float(self._anneal), dtype=torch.float32, device=device)
        return self._anneal_dev

    def step_cb(self, step):
        self._step = step
        self._steps_since_update += 1

    def generate_ray_samples(self, ray_bundle: Optional[RayBundle] = None, density_fns: Optional[List[Callable]] = None,
                             jitters: Optional[List[torch.Tensor]] = None) -> Tuple[RaySamples, List, List]:
        assert ray_bundle is not None and density_fns is not None
        weights_list, ray_samples_list = [], []
        n = self.num_proposal_network_iterations
        weights = ray_samples = None
        updated = self._steps_since_update > self.update_sched(self._step) or self._step < 10
        for i_level in range(n + 1):
            is_prop = i_level < n
            num_samples = self.num_proposal_samples_per_ray[i_level] if is_prop else self.num_nerf_samples_per_ray
            jit = None if jitters is None else jitters[i_level]
            if i_level == 0:
                ray_samples = self.initial_sampler(ray_bundle, num_samples=num_samples, jitter=jit)
            else:
                # the anneal pow (ray_samplers.py:602) is applied inside the resampling kernel
                ray_samples = self.pdf_sampler(ray_bundle, ray_samples, weights, num_samples=num_samples, jitter=jit,
                                               anneal=self._anneal_dev if self._anneal_dev is not None else self._anneal)
            if is_prop:
                fn = density_fns[i_level]
                owner = getattr(fn, "__self__", None)
                # our proposal fields evaluate straight from the ray samples (no positions tensor); any other callable
                # gets the reference's density_fn(positions) call
                fused = getattr(owner, "density_from_ray_samples", None) if getattr(fn, "__name__", "") == "density_fn" else None
                side = None
                if updated and ops.leaf_streams.enabled and i_level < len(ops.leaf_streams.level_streams) and ray_samples.frustums.starts.is_cuda:
                    side = ops.leaf_streams.level_streams[i_level]
                if side is not None:
                    # forward of this level on its own stream (still strictly ordered: fork after, join before, the main stream);
                    # its backward then runs there as well, next to the main field's backward instead of behind it
                    main = torch.cuda.current_stream()
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        density = fused(ray_samples) if fused is not None else fn(ray_samples.frustums.get_positions())
                        weights = ray_samples.get_weights(density)
                    main.wait_stream(side)
                    for t in (density, weights):
                        t.record_stream(main)
                else:
                    with torch.enable_grad() if updated else torch.no_grad():
                        density = fused(ray_samples) if fused is not None else fn(ray_samples.frustums.get_positions())
                    weights = ray_samples.get_weights(density)
                weights_list.append(weights)
                ray_samples_list.append(ray_samples)
        if updated:
            self._steps_since_update = 0
        return ray_samples, weights_list, ray_samples_list
