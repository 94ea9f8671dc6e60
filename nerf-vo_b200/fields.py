"""NerfactoField and HashMLPDensityField with the reference's API
(NS/fields/base_field.py:40-142, nerfacto_field.py:43-297, density_fields.py:34-119)."""
from __future__ import annotations

from enum import Enum
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import ops
from .field_components import MLP, Embedding, HashEncoding, MLPWithHashEncoding, NeRFEncoding, SceneContraction, SHEncoding, repack
from .rays import Frustums, RaySamples


class FieldHeadNames(Enum):
    """NS/field_components/field_heads.py:27-43 (the names the nerfacto path produces)."""

    RGB = "rgb"
    DENSITY = "density"
    NORMALS = "normals"
    PRED_NORMALS = "pred_normals"


class PredNormalsFieldHead(nn.Module):
    """Linear(in_dim, 3) + Tanh + normalize (NS/field_components/field_heads.py:189-204); keeps the `net` key."""

    def __init__(self, in_dim: int) -> None:
        super().__init__()
        self.net = nn.Linear(in_dim, 3)


class Field(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self._sample_locations = None
        self._density_before_activation = None

    def density_fn(self, positions: torch.Tensor, times: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Density at explicit positions [..., 3] -> [..., 1] (NS/fields/base_field.py:48-68)."""
        del times
        density, _ = self._density_from_positions(positions)
        return density

    def get_density(self, ray_samples: RaySamples):
        return self._density_from_positions(ray_samples.frustums.get_positions())

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, torch.Tensor]:
        density, density_embedding = self.get_density(ray_samples)
        field_outputs = self.get_outputs(ray_samples, density_embedding=density_embedding)
        field_outputs[FieldHeadNames.DENSITY] = density
        if compute_normals:
            field_outputs[FieldHeadNames.NORMALS] = self.get_normals()
        return field_outputs


class HashMLPDensityField(Field):
    """Proposal density field (NS/fields/density_fields.py:34-119)."""

    _next_slot = 0

    def __init__(self, aabb: torch.Tensor, num_layers: int = 2, hidden_dim: int = 64, spatial_distortion: Optional[nn.Module] = None,
                 use_linear: bool = False, num_levels: int = 8, max_res: int = 1024, base_res: int = 16, log2_hashmap_size: int = 18,
                 features_per_level: int = 2, implementation: str = "nvo_b200") -> None:
        super().__init__()
        if spatial_distortion is None or not isinstance(spatial_distortion, SceneContraction):
            raise NotImplementedError("nvo_b200 fields require SceneContraction(order=inf) (the NeRF-VO configuration)")
        self.register_buffer("aabb", aabb)
        self.spatial_distortion = spatial_distortion
        self.use_linear = use_linear
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.encoding = HashEncoding(num_levels=num_levels, min_res=base_res, max_res=max_res, log2_hashmap_size=log2_hashmap_size,
                                     features_per_level=features_per_level)
        if not use_linear:
            # density = trunc_exp(mlp(x)) * selector is evaluated inside the MLP kernel (output activation + row mask)
            network = MLP(in_dim=self.encoding.get_out_dim(), num_layers=num_layers, layer_width=hidden_dim, out_dim=1, activation=nn.ReLU(),
                          out_activation="trunc_exp")
            self.mlp_base = nn.Sequential(self.encoding, network)
        else:
            self.linear = MLP(in_dim=self.encoding.get_out_dim(), num_layers=1, layer_width=1, out_dim=1, out_activation="trunc_exp")
        self._fused_ok = None
        # constant-memory bank of the fused kernels: consecutive fields get different banks (proposal network 0 / 1)
        self._slot = HashMLPDensityField._next_slot % 4
        HashMLPDensityField._next_slot += 1

    def _net(self) -> MLP:
        return self.mlp_base[1] if not self.use_linear else self.linear

    def _fused(self) -> bool:
        if self._fused_ok is None:
            self._fused_ok = (not self.use_linear) and ops.prop_density_supported(self.encoding.spec, self._net().spec)
        return self._fused_ok

    def preload(self) -> None:
        """Trainer hook at the start of a step: the MLP goes to its constant-memory bank on a side stream, once, instead of in front of each
        of the step's launches that use it (density forward, backward)."""
        if self._fused():
            net = self._net()
            net._repack()
            ops.prop_density_preload(self._slot, net._flat_param_list())

    def _density_from_positions(self, positions: torch.Tensor) -> Tuple[torch.Tensor, None]:
        net = self._net()
        if self._fused():
            # one kernel: contraction, hash grid, MLP, trunc_exp * selector (csrc/prop.cu)
            net._repack()
            n = positions.numel() // 3
            density = ops.prop_density(self.encoding.hash_table, self.encoding.spec, net.spec, net._flat_param_list(), n, 1, positions=positions, slot=self._slot)
            return density.view(*positions.shape[:-1], 1), None
        x, sel = ops.contract_normalize(positions)
        feat = self.encoding(x)
        density = net(feat, row_mask=sel)
        return density.view(*positions.shape[:-1], 1), None

    def density_from_ray_samples(self, ray_samples: RaySamples) -> torch.Tensor:
        """density_fn(ray_samples.frustums.get_positions()) without materialising the positions: the fused kernel derives
        each sample point from the ray and its interval (same arithmetic as Frustums.get_positions)."""
        fr = ray_samples.frustums
        iv = fr.intervals()
        if not self._fused() or fr.offsets is not None or fr.origins.shape[-2] != 1:
            return self.density_fn(fr.get_positions())
        net = self._net()
        net._repack()
        o = fr.origins.reshape(-1, 3).contiguous()
        d = fr.directions.reshape(-1, 3).contiguous()
        density = ops.prop_density(self.encoding.hash_table, self.encoding.spec, net.spec, net._flat_param_list(), iv.B, iv.S, origins=o, directions=d, iv=iv, slot=self._slot)
        return density.view(iv.B, iv.S, 1)

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[torch.Tensor] = None) -> dict:
        return {}


class NerfactoField(Field):
    """NS/fields/nerfacto_field.py:43-297 (no transient / semantic heads: NeRF-VO does not enable them)."""

    def __init__(self, aabb: torch.Tensor, num_images: int, num_layers: int = 2, hidden_dim: int = 64, geo_feat_dim: int = 15, num_levels: int = 16,
                 base_res: int = 16, max_res: int = 2048, log2_hashmap_size: int = 19, num_layers_color: int = 3, num_layers_transient: int = 2,
                 features_per_level: int = 2, hidden_dim_color: int = 64, hidden_dim_transient: int = 64, appearance_embedding_dim: int = 32,
                 transient_embedding_dim: int = 16, use_transient_embedding: bool = False, use_semantics: bool = False, num_semantic_classes: int = 100,
                 pass_semantic_gradients: bool = False, use_pred_normals: bool = False, use_average_appearance_embedding: bool = False,
                 spatial_distortion: Optional[nn.Module] = None, implementation: str = "nvo_b200", precision: str = "fp16",
                 pred_normals_trainable: bool = True) -> None:
        """precision: "fp16" = hash features and the three MLPs on the tcgen05 tensor-core path (fp16 operands, fp32 accumulate,
        fp32 master parameters — tinycudann's operating point); "fp32" = exact-arithmetic SIMT kernels (bit-level parity runs).
        pred_normals_trainable=False: mlp_pred_normals is evaluated but its activations are not kept for a backward pass (NeRF-VO trains
        with pred_normal_loss_mult = 0, nerf_vo/mapping/nerfstudio.py:74-75); a gradient arriving at pred_normals then raises."""
        super().__init__()
        self.pred_normals_trainable = pred_normals_trainable
        if precision not in ("fp16", "fp32"):
            raise ValueError(f"precision must be 'fp16' or 'fp32', got {precision}")
        self.precision = precision
        if use_transient_embedding or use_semantics:
            raise NotImplementedError("transient / semantic heads are outside the NeRF-VO mapping path")
        if spatial_distortion is None or not isinstance(spatial_distortion, SceneContraction):
            raise NotImplementedError("nvo_b200 fields require SceneContraction(order=inf) (the NeRF-VO configuration)")
        if geo_feat_dim != 15 or appearance_embedding_dim != 32:
            raise NotImplementedError("fused input assembly is specialised for geo_feat_dim=15, appearance_embedding_dim=32 (nerfacto defaults)")
        self.register_buffer("aabb", aabb)
        self.geo_feat_dim = geo_feat_dim
        self.register_buffer("max_res", torch.tensor(max_res))
        self.register_buffer("num_levels", torch.tensor(num_levels))
        self.register_buffer("log2_hashmap_size", torch.tensor(log2_hashmap_size))
        self.spatial_distortion = spatial_distortion
        self.num_images = num_images
        self.appearance_embedding_dim = appearance_embedding_dim
        self.embedding_appearance = Embedding(num_images, appearance_embedding_dim)
        self.use_average_appearance_embedding = use_average_appearance_embedding
        self.use_pred_normals = use_pred_normals
        self.base_res = base_res
        self.direction_encoding = SHEncoding(levels=4)
        self.position_encoding = NeRFEncoding(in_dim=3, num_frequencies=2, min_freq_exp=0, max_freq_exp=1)
        self.mlp_base = MLPWithHashEncoding(num_levels=num_levels, min_res=base_res, max_res=max_res, log2_hashmap_size=log2_hashmap_size,
                                            features_per_level=features_per_level, num_layers=num_layers, layer_width=hidden_dim,
                                            out_dim=1 + geo_feat_dim, activation=nn.ReLU(), out_activation=None)
        if use_pred_normals:
            self.mlp_pred_normals = MLP(in_dim=geo_feat_dim + self.position_encoding.get_out_dim(), num_layers=3, layer_width=64,
                                        out_dim=hidden_dim_transient, activation=nn.ReLU(), out_activation=None)
            self.field_head_pred_normals = PredNormalsFieldHead(in_dim=self.mlp_pred_normals.get_out_dim())
            # mlp_pred_normals (3 layers) + Linear head + Tanh run as ONE 4-layer fused network
            w = 64
            self._pn_spec = ops.MlpSpec(self.mlp_pred_normals.in_dim, (w, w, hidden_dim_transient, 3), acts=("relu", "relu", "none", "tanh"))
        self.mlp_head = MLP(in_dim=self.direction_encoding.get_out_dim() + geo_feat_dim + appearance_embedding_dim, num_layers=num_layers_color,
                            layer_width=hidden_dim_color, out_dim=3, activation=nn.ReLU(), out_activation=nn.Sigmoid())
        self._cache = None
        self._onehot = None

    def tc_networks(self):
        """(parameter list, MlpSpec) of every network the tensor-core path evaluates, in forward order (ops.prepack_weights)."""
        nets = [(self.mlp_base.mlp._flat_param_list(), self.mlp_base.mlp.spec), (self.mlp_head._flat_param_list(), self.mlp_head.spec)]
        if self.use_pred_normals:
            nets.append((self.mlp_pred_normals._flat_param_list() + [self.field_head_pred_normals.net.weight, self.field_head_pred_normals.net.bias],
                         self._pn_spec))
        return nets

    def _pn_params(self):
        if not self.use_pred_normals:
            return []
        return self.mlp_pred_normals._flat_param_list() + [self.field_head_pred_normals.net.weight, self.field_head_pred_normals.net.bias]

    def _fused(self, S: int, on_cuda: bool) -> bool:
        """forward() as the fused kernels of csrc/field_tc.cu (one launch per direction for all three networks)?"""
        return (self.precision == "fp16" and on_cuda and
                ops.field_fused_supported(self.mlp_base.encoder.spec, self.mlp_base.mlp.spec, self.mlp_head.spec,
                                          self._pn_spec if self.use_pred_normals else None, S))

    def prepack(self, S: int) -> None:
        """Trainer hook: pack this step's fp16 weight image(s) on a side stream, ahead of the forward that consumes them."""
        if self.precision != "fp16":
            return
        if self._fused(S, True):
            groups = [self.mlp_base.mlp._flat_param_list(), self.mlp_head._flat_param_list(), self._pn_params()]
            for ps in groups:
                if ps:
                    repack(ps)
            ops.prepack_field_weights(self.mlp_base.encoder.spec, groups[0], groups[1], groups[2] or None)
        else:
            ops.prepack_weights(self.tc_networks())

    def _remember(self, x, h, shape, tc=None) -> None:
        """State get_normals() needs (base_field.py:80-101 keeps _sample_locations / _density_before_activation).  Stored
        DETACHED: normals are first-order and graph-free, and holding autograd nodes across steps would pin the previous
        step's graph (and its stream) — which breaks CUDA-graph capture of the next step.  `tc` = the tensor-core forward's
        internal buffers (features, weight image, saved activations), reused by the input-gradient pass."""
        self._cache = {"x": x.detach(), "shape": tuple(shape), "tc": tc}
        self._sample_locations = self._cache["x"].view(*shape, 3)
        self._density_before_activation = h.detach()[:, :1].view(*shape, 1)

    # -- density -----------------------------------------------------------------------------------------
    def _base(self, x: torch.Tensor, want_normals: bool = False):
        """mlp_base(hash_grid(x)) -> h [n,16]: raw density + geo features (nerfacto_field.py:213-215).  want_normals: the
        grid forward also saves d(feature)/dx (tcnn's dy_dx) so get_normals() needs no second gather pass."""
        enc, mlp = self.mlp_base.encoder, self.mlp_base.mlp
        if self.precision == "fp16":
            mlp._repack()
            tc = {"want_jac": bool(want_normals)}
            return ops.grid_mlp_tc(x, enc.hash_table, enc.spec, mlp.spec, mlp._flat_param_list(), cache=tc), tc
        return mlp(enc(x)), None

    def _density_from_positions(self, positions: torch.Tensor):
        shape = positions.shape[:-1]
        x, sel = ops.contract_normalize(positions)
        h, tc = self._base(x)
        self._remember(x, h, shape, tc)
        density = (ops.trunc_exp(h[:, 0]) * sel).view(*shape, 1)
        return density, h[:, 1:].view(*shape, self.geo_feat_dim)

    def get_normals(self) -> torch.Tensor:
        """-normalize(d raw_density / d x_normalised) (NS/fields/base_field.py:80-101), first order only, no graph."""
        c = self._cache
        assert c is not None, "Sample locations must be set before calling get_normals."
        if c.get("normals") is not None:  # the fused forward evaluated them in the same kernel
            return c["normals"].view(*c["shape"], 3)
        with torch.no_grad():
            enc, mlp = self.mlp_base.encoder, self.mlp_base.mlp
            mlp._repack()
            x = c["x"]
            n = x.shape[0]
            table = enc.hash_table.detach()
            key = (n, mlp.out_dim, str(x.device))
            if self._onehot is None or self._onehot[0] != key:  # constant d(out)/d(out_0) seed, built once per batch shape
                oh = torch.zeros((n, mlp.out_dim), dtype=torch.float32, device=x.device)
                oh[:, 0] = 1.0
                self._onehot = (key, oh)
            onehot = self._onehot[1]
            if self.precision == "fp16":
                tc = c["tc"]  # forward buffers of this very evaluation: only the dgrad chain runs here
                dfeat, _ = ops.mlp_tc_backward(tc["feat16"], tc["wimage"], tc["saved"], tc["y"], onehot, mlp.spec, True, False, dy_absmax=1.0)
            else:
                flat = ops.flat_alias([p.data for p in mlp._flat_param_list()])
                feat = ops.grid_forward(x, table, enc.spec)
                y, saved = ops.mlp_forward(feat, flat, mlp.spec, save=True)
                dfeat, _ = ops.mlp_backward(feat, flat, saved, y, onehot, mlp.spec, need_dx=True, need_dparams=False)
            if self.precision == "fp16" and c["tc"].get("jac") is not None:
                normals = ops.grid_jac_dx(c["tc"]["jac"], dfeat, enc.spec, n, normalize_scale=-1.0, eps=1e-12)
            else:
                g = ops.grid_backward_input(x, table, dfeat, enc.spec, tmf=self.precision == "fp16")
                normals = ops.normalize3(g, scale=-1.0, eps=1e-12)
        return normals.view(*c["shape"], 3)

    # -- colour / predicted normals -------------------------------------------------------------------------
    def _appearance(self, ray_samples: RaySamples, B: int, device):
        """(camera indices [B] | None, embedding table | mean vector): nerfacto_field.py:236-251."""
        if self.training:
            return ray_samples.camera_indices.reshape(B).long().contiguous(), self.embedding_appearance.embedding.weight
        if self.use_average_appearance_embedding:
            return None, self.embedding_appearance.mean(dim=0)
        return None, torch.zeros(self.appearance_embedding_dim, device=device)

    def _heads(self, h, sel, positions, dirs, cam, emb, B: int, S: int) -> Dict[FieldHeadNames, torch.Tensor]:
        """Input assembly (SH | geo | appearance, posenc | geo) -> mlp_head -> rgb, mlp_pred_normals + head -> pred_normals, and
        density = trunc_exp(h0) * selector, from the base network's output h [n,16] (nerfacto_field.py:225-297)."""
        out: Dict[FieldHeadNames, torch.Tensor] = {}
        pn_params = ()
        if self.use_pred_normals:
            pn_params = self.mlp_pred_normals._flat_param_list() + [self.field_head_pred_normals.net.weight, self.field_head_pred_normals.net.bias]
            repack(pn_params)
        if self.precision == "fp16":
            self.mlp_head._repack()
            density, rgb, pn = ops.field_heads_tc(h, emb, sel, dirs, positions.reshape(-1, 3), cam, B, S, self.mlp_head.spec,
                                                  self.mlp_head._flat_param_list(), self._pn_spec if self.use_pred_normals else None, pn_params)
            if pn is not None:
                out[FieldHeadNames.PRED_NORMALS] = pn.view(B, S, 3)
            out[FieldHeadNames.RGB] = rgb.view(B, S, 3)
            out[FieldHeadNames.DENSITY] = density.view(B, S, 1)
            return out
        density, head_in, pn_in = ops.field_assemble(h, emb, sel, dirs, positions.reshape(-1, 3), cam, B, S, self.use_pred_normals)
        if self.use_pred_normals:
            pn = ops.mlp_apply(pn_in, self._pn_spec, pn_params)
            out[FieldHeadNames.PRED_NORMALS] = ops.normalize3(pn, 1.0, 1e-12).view(B, S, 3)
        out[FieldHeadNames.RGB] = self.mlp_head(head_in).view(B, S, 3)
        out[FieldHeadNames.DENSITY] = density.view(B, S, 1)
        return out

    def get_outputs(self, ray_samples: RaySamples, density_embedding: Optional[torch.Tensor] = None) -> Dict[FieldHeadNames, torch.Tensor]:
        """NerfactoField.get_outputs (NS/fields/nerfacto_field.py:225-297; contract NS/fields/base_field.py:104-112): colour and predicted
        normals from the ray samples and the geometry features get_density() returned.  The base class calls it right after get_density
        (base_field.py:114-133); forward() below runs the same kernels as one pipeline that also produces the density."""
        assert density_embedding is not None
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")
        fr = ray_samples.frustums
        B, S = fr.shape
        n = B * S
        geo = density_embedding.reshape(n, self.geo_feat_dim)
        # h = [raw density | geo]: the raw-density column only feeds the density output, which this call does not return
        h = torch.cat([torch.zeros((n, 1), dtype=geo.dtype, device=geo.device), geo], dim=1)
        sel = torch.ones(n, dtype=torch.float32, device=geo.device)
        cam, emb = self._appearance(ray_samples, B, geo.device)
        out = self._heads(h, sel, fr.get_positions(), fr.directions.reshape(B, 3).contiguous(), cam, emb, B, S)
        out.pop(FieldHeadNames.DENSITY)
        return out

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False) -> Dict[FieldHeadNames, torch.Tensor]:
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")
        fr = ray_samples.frustums
        B, S = fr.shape
        per_ray = fr.offsets is None and fr.origins.dim() == 3 and fr.origins.shape[-2] == 1
        if per_ray and self._fused(S, fr.origins.is_cuda):
            enc = self.mlp_base.encoder
            groups = [self.mlp_base.mlp._flat_param_list(), self.mlp_head._flat_param_list(), self._pn_params()]
            for ps in groups:
                if ps:
                    repack(ps)
            o, dirs = fr.origins.reshape(B, 3), fr.directions.reshape(B, 3)
            cam, emb = self._appearance(ray_samples, B, dirs.device)
            density, rgb, pn, normals, h0, x = ops.field_fused(
                o, dirs, fr.intervals(), enc.hash_table, emb, cam, B, S, enc.spec, compute_normals, groups[0], groups[1],
                self._pn_spec if self.use_pred_normals else None, groups[2], save_pn=self.pred_normals_trainable)
            self._cache = {"x": x, "shape": (B, S), "tc": None, "normals": normals}
            self._sample_locations = x.view(B, S, 3)
            self._density_before_activation = h0.view(B, S, 1)
            out = {FieldHeadNames.RGB: rgb.view(B, S, 3), FieldHeadNames.DENSITY: density.view(B, S, 1)}
            if pn is not None:
                out[FieldHeadNames.PRED_NORMALS] = pn.view(B, S, 3)
            if compute_normals:
                out[FieldHeadNames.NORMALS] = normals.view(B, S, 3)
            return out
        positions = fr.get_positions()
        x, sel = ops.contract_normalize(positions)
        dirs = fr.directions.reshape(B, 3).contiguous()
        cam, emb = self._appearance(ray_samples, B, dirs.device)
        h, tc = self._base(x, want_normals=compute_normals)
        self._remember(x, h, (B, S), tc)
        normals = side_n = None
        if self.precision == "fp16" and compute_normals and ops.leaf_streams.enabled and x.is_cuda and ops.env_flag("NVO_FIELD_BRANCHES", True):
            # density-gradient normals (base network's input-gradient chain + saved-Jacobian product) depend on the base network only:
            # they run on their own stream next to the colour / predicted-normals heads and are joined before the renderer reads them
            main = torch.cuda.current_stream()
            side_n = ops.leaf_streams.branch_streams[0]
            side_n.wait_stream(main)
            with torch.cuda.stream(side_n):
                normals = self.get_normals()
        out = self._heads(h, sel, positions, dirs, cam, emb, B, S)
        if side_n is not None:
            main.wait_stream(side_n)
            normals.record_stream(main)
            out[FieldHeadNames.NORMALS] = normals
        elif compute_normals:
            out[FieldHeadNames.NORMALS] = self.get_normals()
        return out
