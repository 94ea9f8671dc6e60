"""nerfstudio field components with the reference's constructor signatures and state-dict keys
(NS/field_components/{encodings,mlp,activations,spatial_distortions,embedding,field_heads}.py), computing through
the nvo_b200 kernels.  `implementation` is accepted for signature compatibility; there is one implementation."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import ops

trunc_exp = ops.trunc_exp


def sh_encode(x: torch.Tensor) -> torch.Tensor:
    """degree-4 SH of the vector as given, no gradient (NS/field_components/encodings.py:796-799)."""
    with torch.no_grad():
        return ops.sh4(x).view(*x.shape[:-1], 16)


def frequency_encode(x: torch.Tensor, n_freq: int) -> torch.Tensor:
    return ops.frequency(x.reshape(-1, x.shape[-1]).contiguous().float(), n_freq).view(*x.shape[:-1], -1)


def _activation_name(act) -> str:
    if act is None:
        return "none"
    if isinstance(act, str):
        return act.lower()
    name = type(act).__name__.lower()
    table = {"relu": "relu", "sigmoid": "sigmoid", "tanh": "tanh", "identity": "none"}
    if name not in table:
        raise ValueError(f"activation {act} not supported")
    return table[name]


def repack(params) -> None:
    """Re-home `params` as back-to-back views of ONE flat fp32 buffer (no-op when they already are), so the kernels
    take a single pointer (the tcnn `params` layout) while state-dict keys stay those of the reference's torch modules."""
    if ops.flat_alias([p.data for p in params]) is not None:
        return
    flat = torch.cat([p.data.reshape(-1).float() for p in params])
    off = 0
    for p in params:
        n = p.numel()
        p.data = flat[off:off + n].view(p.shape)
        off += n


class FlatParamsMixin:
    def _flat_param_list(self):
        raise NotImplementedError

    def _repack(self):
        repack(self._flat_param_list())


class HashEncoding(nn.Module):
    """NS/field_components/encodings.py:314-470."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19, features_per_level: int = 2,
                 hash_init_scale: float = 0.001, implementation: str = "nvo_b200", interpolation: Optional[str] = None) -> None:
        super().__init__()
        if interpolation not in (None, "Linear"):
            raise AssertionError(f"interpolation '{interpolation}' is not supported")
        self.in_dim = 3
        self.num_levels, self.min_res, self.max_res = num_levels, min_res, max_res
        self.features_per_level = features_per_level
        self.hash_init_scale = hash_init_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.hash_table_size = 2**log2_hashmap_size
        self.growth_factor = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
        self.scalings = ops.torch_level_scalings(num_levels, min_res, max_res)
        self.hash_offset = torch.arange(num_levels) * self.hash_table_size
        self.spec = ops.GridSpec(num_levels, log2_hashmap_size, tuple(float(s) for s in self.scalings), features_per_level)
        table = (torch.rand(size=(self.hash_table_size * num_levels, features_per_level)) * 2 - 1) * hash_init_scale
        self.hash_table = nn.Parameter(table)

    def get_out_dim(self) -> int:
        return self.num_levels * self.features_per_level

    def hash_indices(self, in_tensor: torch.Tensor) -> torch.Tensor:
        """Rows of all 8 corners, [N, L, 8] int64, reference corner order (test hook for hash_fn :405-422)."""
        return ops.grid_indices(in_tensor.reshape(-1, 3).contiguous().float(), self.spec)

    def forward(self, in_tensor: torch.Tensor) -> torch.Tensor:
        x = in_tensor.reshape(-1, 3)
        y = ops.grid_encode(x.float(), self.hash_table, self.spec)
        return y.view(*in_tensor.shape[:-1], -1)


class MLP(nn.Module, FlatParamsMixin):
    """NS/field_components/mlp.py:61-185 (skip connections unsupported: nerfacto does not use them)."""

    def __init__(self, in_dim: int, num_layers: int, layer_width: int, out_dim: Optional[int] = None, skip_connections: Optional[Tuple[int]] = None,
                 activation=nn.ReLU(), out_activation=None, implementation: str = "nvo_b200") -> None:
        super().__init__()
        assert in_dim > 0
        if skip_connections:
            raise NotImplementedError("skip connections are not supported by nvo_b200.MLP")
        self.in_dim = in_dim
        self.out_dim = out_dim if out_dim is not None else layer_width
        self.num_layers, self.layer_width = num_layers, layer_width
        self.activation, self.out_activation = activation, out_activation
        dims = [layer_width] * (num_layers - 1) + [self.out_dim]
        self.layers = nn.ModuleList([nn.Linear(i, o) for i, o in zip([in_dim] + dims[:-1], dims)])
        self.spec = ops.MlpSpec(in_dim, tuple(dims), _activation_name(activation), _activation_name(out_activation))

    def get_out_dim(self) -> int:
        return self.out_dim

    def _flat_param_list(self):
        out = []
        for layer in self.layers:
            out += [layer.weight, layer.bias]
        return out

    def forward(self, in_tensor: torch.Tensor, row_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._repack()
        x = in_tensor.reshape(-1, self.in_dim).float()
        y = ops.mlp_apply(x, self.spec, self._flat_param_list(), row_mask)
        return y.view(*in_tensor.shape[:-1], self.out_dim)


class MLPWithHashEncoding(nn.Module):
    """NS/field_components/mlp.py:187-295; state-dict keys `model.0.hash_table`, `model.1.layers.i.{weight,bias}`."""

    def __init__(self, num_levels: int = 16, min_res: int = 16, max_res: int = 1024, log2_hashmap_size: int = 19, features_per_level: int = 2,
                 hash_init_scale: float = 0.001, interpolation: Optional[str] = None, num_layers: int = 2, layer_width: int = 64,
                 out_dim: Optional[int] = None, skip_connections=None, activation=nn.ReLU(), out_activation=None, implementation: str = "nvo_b200") -> None:
        super().__init__()
        self.in_dim = 3
        self.out_dim = out_dim if out_dim is not None else layer_width
        enc = HashEncoding(num_levels, min_res, max_res, log2_hashmap_size, features_per_level, hash_init_scale, interpolation=interpolation)
        mlp = MLP(enc.get_out_dim(), num_layers, layer_width, self.out_dim, skip_connections, activation, out_activation)
        self.model = nn.Sequential(enc, mlp)

    @property
    def encoder(self) -> HashEncoding:
        return self.model[0]

    @property
    def mlp(self) -> MLP:
        return self.model[1]

    def get_out_dim(self) -> int:
        return self.out_dim

    def forward(self, in_tensor: torch.Tensor) -> torch.Tensor:
        return self.model(in_tensor)


class SHEncoding(nn.Module):
    """NS/field_components/encodings.py:759-804."""

    def __init__(self, levels: int = 4, implementation: str = "nvo_b200") -> None:
        super().__init__()
        if levels != 4:
            raise ValueError(f"nvo_b200 SHEncoding supports levels == 4 (nerfacto), requested {levels}")
        self.levels = levels
        self.in_dim = 3

    def get_out_dim(self) -> int:
        return self.levels**2

    def forward(self, in_tensor: torch.Tensor) -> torch.Tensor:
        return sh_encode(in_tensor.float())


class NeRFEncoding(nn.Module):
    """NS/field_components/encodings.py:102-194 (no covariances, power-of-two frequencies)."""

    def __init__(self, in_dim: int, num_frequencies: int, min_freq_exp: float, max_freq_exp: float, include_input: bool = False,
                 implementation: str = "nvo_b200") -> None:
        super().__init__()
        if min_freq_exp != 0 or max_freq_exp != num_frequencies - 1:
            raise ValueError("nvo_b200 NeRFEncoding needs min_freq_exp == 0 and max_freq_exp == num_frequencies - 1")
        self.in_dim, self.num_frequencies, self.include_input = in_dim, num_frequencies, include_input

    def get_out_dim(self) -> int:
        return self.in_dim * self.num_frequencies * 2 + (self.in_dim if self.include_input else 0)

    def forward(self, in_tensor: torch.Tensor, covs=None) -> torch.Tensor:
        if covs is not None:
            raise NotImplementedError("integrated positional encoding is not on the nerfacto path")
        enc = frequency_encode(in_tensor, self.num_frequencies)
        return torch.cat([enc, in_tensor], dim=-1) if self.include_input else enc


class SceneContraction(nn.Module):
    """L-infinity contraction (NS/field_components/spatial_distortions.py:42-90); order must be inf."""

    def __init__(self, order=float("inf")) -> None:
        super().__init__()
        if order != float("inf"):
            raise NotImplementedError("nvo_b200 SceneContraction implements order=inf (what nerfacto uses)")
        self.order = order

    def forward(self, positions: torch.Tensor) -> torch.Tensor:
        """Returns the contracted positions (in [-2,2]^3)."""
        x, _ = self.normalized(positions, mask=False)
        return x.view(positions.shape) * 4.0 - 2.0

    @staticmethod
    def normalized(positions: torch.Tensor, mask: bool = True):
        return ops.contract_normalize(positions)


class Embedding(nn.Module):
    """NS/field_components/embedding.py:27-55."""

    def __init__(self, in_dim: int, out_dim: int) -> None:
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.embedding = nn.Embedding(in_dim, out_dim)

    def mean(self, dim=0):
        return self.embedding.weight.mean(dim)

    def forward(self, in_tensor: torch.Tensor) -> torch.Tensor:
        return self.embedding(in_tensor)
