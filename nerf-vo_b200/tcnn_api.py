"""tinycudann-compatible modules (`Encoding`, `Network`, `NetworkWithInputEncoding`) backed by the nvo_b200
kernels.  Mirrors nerf_vo/thirdparty/tiny_cuda_nn/bindings/torch/tinycudann/modules.py:162-329: same constructor
arguments (JSON-style config dicts), a single flat fp32 `params` Parameter, `n_input_dims`, `n_output_dims`,
`dtype`, `seed`, `loss_scale`, `forward(x)` on [N, n_in] inputs, pickling support.

Semantic convention: the arithmetic is the reference's *torch* implementation (the parity oracle; SURVEY.md §8
a-notes): every level hashed into a 2^k slab, floor/ceil corners, biased Linear layers in torch layout, grid init
U(+-1e-3).  Outputs are fp32 (`dtype=torch.float32`); padding the batch to `batch_size_granularity()` is not needed.
Drop-in use:   `import nerf_vo_b200.tcnn_api as tcnn`   (cf. NS/utils/external.py:38-58).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from . import _lib, ops


def batch_size_granularity() -> int:
    return int(_lib.load().nvo_batch_size_granularity())


def free_temporary_memory() -> None:
    """bindings.cpp:287 — the library owns no memory; torch's caching allocator does."""
    torch.cuda.empty_cache()


def has_networks() -> bool:
    return True


def preferred_precision():
    return torch.float32


def _act(name) -> str:
    name = (name or "None").lower()
    if name not in _lib.ACT:
        raise RuntimeError(f"Invalid activation type: {name}")
    return name


def _linear_init_(w: torch.Tensor, b: torch.Tensor, gen: torch.Generator) -> None:
    bound = 1.0 / math.sqrt(w.shape[1])
    w.copy_((torch.rand(w.shape, generator=gen) * 2 - 1) * bound)
    b.copy_((torch.rand(b.shape, generator=gen) * 2 - 1) * bound)


class Module(torch.nn.Module):
    """Base: one flat fp32 `params` tensor, like tinycudann.Module (modules.py:162-207)."""

    def __init__(self, seed: int = 1337):
        super().__init__()
        self.seed = seed
        self.dtype = torch.float32
        self.loss_scale = 1.0  # fp32 gradients: no loss scaling needed (modules.py:174 uses 128 for fp16)
        self.params = torch.nn.Parameter(self._initial_params(seed), requires_grad=True)

    def _initial_params(self, seed: int) -> torch.Tensor:
        raise NotImplementedError

    def _forward(self, x: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("input must be a CUDA tensor (nvo_b200 has no CPU path)")
        if x.dim() != 2 or x.shape[1] != self.n_input_dims:
            raise RuntimeError(f"input must have shape [N, {self.n_input_dims}], got {tuple(x.shape)}")
        if not self.params.is_cuda:
            raise RuntimeError("module parameters must be on a CUDA device: call .cuda() first")
        return self._forward(x.to(torch.float32).contiguous())[:, : self.n_output_dims]

    def extra_repr(self) -> str:
        return f"n_input_dims={self.n_input_dims}, n_output_dims={self.n_output_dims}, seed={self.seed}, dtype={self.dtype}, hyperparams={self.hyperparams()}"


class _GridPart:
    def __init__(self, n_input_dims: int, cfg: Dict):
        if n_input_dims != 3:
            raise RuntimeError("HashGrid encoding supports n_input_dims == 3")
        interp = cfg.get("interpolation", "Linear")
        if interp != "Linear":
            raise RuntimeError(f"interpolation '{interp}' is not supported (Linear only, as the reference torch path)")
        L = int(cfg.get("n_levels", 16))
        F = int(cfg.get("n_features_per_level", 2))
        self.log2_T = int(cfg.get("log2_hashmap_size", 19))
        base = int(cfg.get("base_resolution", 16))
        pls = float(cfg.get("per_level_scale", 2.0))
        sc = ops.growth_level_scalings(L, base, pls)
        self.spec = ops.GridSpec(L, self.log2_T, tuple(float(s) for s in sc), F)
        self.n_params = self.spec.n_rows * F
        self.n_output_dims = L * F
        self.init_scale = float(cfg.get("hash_init_scale", 1e-3))
        self.cfg = dict(cfg)

    def initial_params(self, gen) -> torch.Tensor:
        return (torch.rand(self.n_params, generator=gen) * 2 - 1) * self.init_scale

    def apply(self, x, flat):
        return ops.grid_encode(x, flat, self.spec)


class _MlpPart:
    def __init__(self, n_input_dims: int, n_output_dims: int, cfg: Dict):
        otype = cfg.get("otype", "FullyFusedMLP")
        if otype not in ("FullyFusedMLP", "CutlassMLP"):
            raise RuntimeError(f"Invalid network type: {otype}")
        width = int(cfg.get("n_neurons", 64))
        hidden = int(cfg.get("n_hidden_layers", 1))
        dims = [width] * hidden + [int(n_output_dims)]
        self.spec = ops.MlpSpec(int(n_input_dims), tuple(dims), _act(cfg.get("activation", "ReLU")), _act(cfg.get("output_activation", "None")))
        self.n_params = self.spec.n_params
        self.n_output_dims = int(n_output_dims)
        self.cfg = dict(cfg)

    def initial_params(self, gen) -> torch.Tensor:
        flat = torch.empty(self.n_params)
        for (a, b), (ws, bs) in zip(ops._pairs(self.spec.offsets()), self.spec.shapes):
            _linear_init_(flat[a[0]:a[1]].view(ws), flat[b[0]:b[1]].view(bs), gen)
        return flat

    def apply(self, x, flat):
        views: List[torch.Tensor] = []
        for (a, b), (ws, bs) in zip(ops._pairs(self.spec.offsets()), self.spec.shapes):
            views.append(flat[a[0]:a[1]].view(ws))
            views.append(flat[b[0]:b[1]].view(bs))
        if ops.tc_eligible(self.spec) and min(self.spec.dims[:-1] or (64,)) >= 32:
            return ops.mlp_apply_tc(x, self.spec, views)  # FullyFusedMLP on tcgen05 (fp16 operands, fp32 accumulate)
        return ops.mlp_apply(x, self.spec, views)  # narrow networks: exact fp32 SIMT kernel


class Encoding(Module):
    """tcnn.Encoding(n_input_dims, encoding_config, seed=1337, dtype=None) — modules.py:286-329.
    Supported otypes: HashGrid, SphericalHarmonics (degree 4), Frequency."""

    def __init__(self, n_input_dims: int, encoding_config: Dict, seed: int = 1337, dtype=None):
        self.n_input_dims = int(n_input_dims)
        self.encoding_config = dict(encoding_config)
        otype = self.encoding_config.get("otype", "HashGrid")
        self._otype = otype
        if otype in ("HashGrid", "Grid"):
            self._grid = _GridPart(self.n_input_dims, self.encoding_config)
            self.n_output_dims = self._grid.n_output_dims
        elif otype == "SphericalHarmonics":
            if int(self.encoding_config.get("degree", 4)) != 4 or self.n_input_dims != 3:
                raise RuntimeError("SphericalHarmonics: only degree 4 on 3 input dims is supported")
            self.n_output_dims = 16
        elif otype == "Frequency":
            self._n_freq = int(self.encoding_config.get("n_frequencies", 2))
            self.n_output_dims = self.n_input_dims * self._n_freq * 2
        else:
            raise RuntimeError(f"Invalid encoding type: {otype}")
        super().__init__(seed)
        if dtype is not None and dtype != torch.float32:
            raise RuntimeError("nvo_b200 Encoding outputs float32")

    def _initial_params(self, seed):
        if self._otype in ("HashGrid", "Grid"):
            return self._grid.initial_params(torch.Generator().manual_seed(seed))
        return torch.zeros(0)

    def hyperparams(self):
        return dict(self.encoding_config)

    def _forward(self, x):
        from . import field_components as fc

        if self._otype in ("HashGrid", "Grid"):
            return self._grid.apply(x, self.params)
        if self._otype == "SphericalHarmonics":
            return fc.sh_encode(x)
        return fc.frequency_encode(x, self._n_freq)


class Network(Module):
    """tcnn.Network(n_input_dims, n_output_dims, network_config, seed=1337) — modules.py:251-284."""

    def __init__(self, n_input_dims: int, n_output_dims: int, network_config: Dict, seed: int = 1337):
        self.n_input_dims = int(n_input_dims)
        self.n_output_dims = int(n_output_dims)
        self.network_config = dict(network_config)
        self._mlp = _MlpPart(self.n_input_dims, self.n_output_dims, self.network_config)
        super().__init__(seed)

    def _initial_params(self, seed):
        return self._mlp.initial_params(torch.Generator().manual_seed(seed))

    def hyperparams(self):
        return dict(self.network_config)

    def _forward(self, x):
        return self._mlp.apply(x, self.params)


class NetworkWithInputEncoding(Module):
    """tcnn.NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config, network_config, seed) —
    modules.py:209-249.  Parameter layout = [network | encoding] (network_with_input_encoding.h:115-134)."""

    def __init__(self, n_input_dims: int, n_output_dims: int, encoding_config: Dict, network_config: Dict, seed: int = 1337):
        self.n_input_dims = int(n_input_dims)
        self.n_output_dims = int(n_output_dims)
        self.encoding_config = dict(encoding_config)
        self.network_config = dict(network_config)
        if self.encoding_config.get("otype", "HashGrid") not in ("HashGrid", "Grid"):
            raise RuntimeError("NetworkWithInputEncoding supports a HashGrid input encoding")
        self._grid = _GridPart(self.n_input_dims, self.encoding_config)
        self._mlp = _MlpPart(self._grid.n_output_dims, self.n_output_dims, self.network_config)
        super().__init__(seed)

    def _initial_params(self, seed):
        gen = torch.Generator().manual_seed(seed)
        return torch.cat([self._mlp.initial_params(gen), self._grid.initial_params(gen)])

    def hyperparams(self):
        return {"encoding": dict(self.encoding_config), "network": dict(self.network_config)}

    def _forward(self, x):
        n = self._mlp.n_params
        feat = self._grid.apply(x, self.params[n:])
        return self._mlp.apply(feat, self.params[:n])
