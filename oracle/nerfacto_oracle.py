"""CPU ORACLE — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A from-scratch CPU restatement (torch CPU tensors used as a numpy-with-autograd; fp32 unless
noted) of the algorithm of NeRF-VO's mapping hot path, i.e. the nerfstudio *torch*
implementation that `implementation="torch"` selects in the reference.  Every function cites
the reference file:line it follows (paths relative to /root/reference, `NS/` =
nerf_vo/thirdparty/nerfstudio/nerfstudio/).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker or as the timed CPU baseline.  The
product (`nerf-vo_b200/`) never imports it and has no CPU fallback.

Parity pinning: this restatement is checked against (a) the unmodified reference executed in
the build container (`tests/test_oracle_vs_reference.py`, skipped where /root/reference is
absent) and (b) golden vectors generated from the reference by `oracle/make_golden.py`
and committed under `tests/golden/` (`tests/test_oracle_golden.py`, runs everywhere).

Design notes (why torch-CPU and not numpy): the path needs parameter gradients; autograd on
the restated forward gives them without a second hand-derived implementation that could share
a mistake with the CUDA backward.  Integer work (hash indices, searchsorted) is int64/int32
exactly as the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------------------
# configuration (defaults = NeRF-VO's; NS/models/nerfacto.py:55-131, nerf_vo/mapping/nerfstudio.py:47-83)
# ----------------------------------------------------------------------------------------


@dataclass
class GridCfg:
    num_levels: int = 16
    min_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2


@dataclass
class ModelCfg:
    main_grid: GridCfg = field(default_factory=GridCfg)
    prop_grids: Tuple[GridCfg, ...] = (
        GridCfg(num_levels=5, min_res=16, max_res=128, log2_hashmap_size=17),
        GridCfg(num_levels=5, min_res=16, max_res=256, log2_hashmap_size=17),
    )
    num_images: int = 192
    hidden_dim: int = 64
    geo_feat_dim: int = 15
    appearance_dim: int = 32
    prop_hidden_dim: int = 16
    num_proposal_samples: Tuple[int, ...] = (256, 96)
    num_nerf_samples: int = 48
    near: float = 0.05
    far: float = 1000.0
    predict_normals: bool = True
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    orientation_loss_mult: float = 0.0
    pred_normal_loss_mult: float = 0.0
    depth_loss_mult: float = 0.001
    normal_loss_mult: float = 0.000005
    depth_sigma: float = 0.001
    histogram_padding: float = 0.01


# ----------------------------------------------------------------------------------------
# a1  hash grid  (NS/field_components/encodings.py:328-465)
# ----------------------------------------------------------------------------------------

HASH_PRIMES = (1, 2654435761, 805459861)  # encodings.py:417


def level_scalings(cfg: GridCfg) -> Tensor:
    """fp32 per-level scale, evaluated exactly as encodings.py:347-349 does (float32 pow, floor)."""
    levels = torch.arange(cfg.num_levels)
    growth = np.exp((np.log(cfg.max_res) - np.log(cfg.min_res)) / (cfg.num_levels - 1)) if cfg.num_levels > 1 else 1
    return torch.floor(cfg.min_res * growth**levels)


def grid_corners(x: Tensor, scalings: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """scaled / ceil / floor integer coordinates and fractional offset (encodings.py:428-433).

    x [N,3] fp32 -> (c [N,L,3] int32, f [N,L,3] int32, offset [N,L,3] fp32)
    """
    scaled = x[..., None, :] * scalings.view(-1, 1).to(x)
    c = torch.ceil(scaled).to(torch.int32)
    f = torch.floor(scaled).to(torch.int32)
    return c, f, scaled - f


# corner k of the reference picks ceil (1) or floor (0) per axis (encodings.py:435-442)
CORNER_SELECT = ((1, 1, 1), (1, 0, 1), (0, 0, 1), (0, 1, 1), (1, 1, 0), (1, 0, 0), (0, 0, 0), (0, 1, 0))


def hash_coords(v: Tensor, log2_T: int) -> Tensor:
    """Spatial hash of int coords [...,L,3] into the level's slab, in int64 (encodings.py:417-421)."""
    L = v.shape[-2]
    v = v.to(torch.int64)
    h = (v[..., 0] * HASH_PRIMES[0]) ^ (v[..., 1] * HASH_PRIMES[1]) ^ (v[..., 2] * HASH_PRIMES[2])
    h = h % (1 << log2_T)
    return h + torch.arange(L, dtype=torch.int64) * (1 << log2_T)


def hash_indices(x: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    """All 8 corner row indices, reference corner order: int64 [N, L, 8]."""
    c, f, _ = grid_corners(x, scalings)
    out = []
    for sel in CORNER_SELECT:
        v = torch.stack([c[..., a] if sel[a] else f[..., a] for a in range(3)], dim=-1)
        out.append(hash_coords(v, log2_T))
    return torch.stack(out, dim=-1)


def hash_encode(x: Tensor, table: Tensor, scalings: Tensor, log2_T: int) -> Tensor:
    """Trilinear hash-grid features [N, L*F]; interpolation order x, y, z (encodings.py:444-465)."""
    _, _, o = grid_corners(x, scalings)
    idx = hash_indices(x, scalings, log2_T)
    fe = [table[idx[..., k]] for k in range(8)]  # each [N,L,F]
    ox, oy, oz = o[..., 0:1], o[..., 1:2], o[..., 2:3]
    f03 = fe[0] * ox + fe[3] * (1 - ox)
    f12 = fe[1] * ox + fe[2] * (1 - ox)
    f56 = fe[5] * ox + fe[6] * (1 - ox)
    f47 = fe[4] * ox + fe[7] * (1 - ox)
    f0312 = f03 * oy + f12 * (1 - oy)
    f4756 = f47 * oy + f56 * (1 - oy)
    enc = f0312 * oz + f4756 * (1 - oz)
    return enc.flatten(-2)


# ----------------------------------------------------------------------------------------
# a3  MLP  (NS/field_components/mlp.py:143-179)
# ----------------------------------------------------------------------------------------


def mlp_forward(x: Tensor, weights: Sequence[Tensor], biases: Sequence[Tensor], out_activation: Optional[str] = None) -> Tensor:
    """y = W x + b, ReLU between layers, optional output activation ('sigmoid'|'relu'|None)."""
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        x = x @ w.t() + b
        if i < n - 1:
            x = torch.relu(x)
    if out_activation == "sigmoid":
        x = torch.sigmoid(x)
    elif out_activation == "relu":
        x = torch.relu(x)
    elif out_activation not in (None, "none"):
        raise ValueError(out_activation)
    return x


# ----------------------------------------------------------------------------------------
# a5  direction / position encodings
# ----------------------------------------------------------------------------------------


def sh_deg4(d: Tensor) -> Tensor:
    """16 real SH components of the vector AS GIVEN (NS/utils/math.py:45-78; no remap, no grad)."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x**2, y**2, z**2
    c = torch.zeros((*d.shape[:-1], 16), dtype=d.dtype)
    c[..., 0] = 0.28209479177387814
    c[..., 1] = 0.4886025119029199 * y
    c[..., 2] = 0.4886025119029199 * z
    c[..., 3] = 0.4886025119029199 * x
    c[..., 4] = 1.0925484305920792 * x * y
    c[..., 5] = 1.0925484305920792 * y * z
    c[..., 6] = 0.9461746957575601 * zz - 0.31539156525251999
    c[..., 7] = 1.0925484305920792 * x * z
    c[..., 8] = 0.5462742152960396 * (xx - yy)
    c[..., 9] = 0.5900435899266435 * y * (3 * xx - yy)
    c[..., 10] = 2.890611442640554 * x * y * z
    c[..., 11] = 0.4570457994644658 * y * (5 * zz - 1)
    c[..., 12] = 0.3731763325901154 * z * (5 * zz - 3)
    c[..., 13] = 0.4570457994644658 * x * (5 * zz - 1)
    c[..., 14] = 1.445305721320277 * z * (xx - yy)
    c[..., 15] = 0.5900435899266435 * x * (xx - 3 * yy)
    return c.detach()


def posenc_2freq(p: Tensor) -> Tensor:
    """sin(cat[u, u + pi/2]), u = (2*pi*p)[..., None] * [1, 2] flattened (encodings.py:170-176)."""
    s = 2 * torch.pi * p
    freqs = 2 ** torch.linspace(0.0, 1.0, 2)
    u = (s[..., None] * freqs).reshape(*s.shape[:-1], -1)
    return torch.sin(torch.cat([u, u + torch.pi / 2.0], dim=-1))


# ----------------------------------------------------------------------------------------
# a6 / a7  contraction, trunc_exp
# ----------------------------------------------------------------------------------------


def contract_linf(p: Tensor) -> Tensor:
    """L-infinity scene contraction (NS/field_components/spatial_distortions.py:67-69)."""
    mag = torch.linalg.norm(p, ord=float("inf"), dim=-1)[..., None]
    return torch.where(mag < 1, p, (2 - (1 / mag)) * (p / mag))


def normalized_positions(p: Tensor) -> Tuple[Tensor, Tensor]:
    """(contract(p)+2)/4, zeroed where any coordinate leaves (0,1) (NS/fields/nerfacto_field.py:201-209)."""
    x = (contract_linf(p) + 2.0) / 4.0
    sel = ((x > 0.0) & (x < 1.0)).all(dim=-1)
    return x * sel[..., None], sel


class _TruncExp(torch.autograd.Function):
    """exp forward; backward g*exp(clamp(x,-15,15)) (NS/field_components/activations.py:28-41)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


trunc_exp = _TruncExp.apply


# ----------------------------------------------------------------------------------------
# parameters (keys = the reference state_dict keys of DepthNerfactoModel, torch implementation)
# ----------------------------------------------------------------------------------------


def _linear_init(gen: torch.Generator, out_dim: int, in_dim: int) -> Tuple[Tensor, Tensor]:
    """nn.Linear default init: U(+-1/sqrt(in)) for weight (kaiming a=sqrt5) and bias."""
    bound = 1.0 / math.sqrt(in_dim)
    w = (torch.rand(out_dim, in_dim, generator=gen) * 2 - 1) * bound
    b = (torch.rand(out_dim, generator=gen) * 2 - 1) * bound
    return w, b


def init_params(cfg: ModelCfg, seed: int = 0, table_std: Optional[float] = None) -> Dict[str, Tensor]:
    """Random-init parameters with the reference's shapes/keys and init distributions
    (hash table U(+-1e-3) encodings.py:381-382; nn.Linear default; nn.Embedding N(0,1)).
    `table_std` switches the tables to N(0, table_std) — the 'trained-like' state of SURVEY §8d.
    Not bit-identical to the reference's RNG stream; parity tests always pass parameters explicitly.
    """
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, Tensor] = {}

    def table(gc: GridCfg) -> Tensor:
        n = (1 << gc.log2_hashmap_size) * gc.num_levels
        if table_std is None:
            return (torch.rand(n, gc.features_per_level, generator=g) * 2 - 1) * 1e-3
        return torch.randn(n, gc.features_per_level, generator=g) * table_std

    P["field.embedding_appearance.embedding.weight"] = torch.randn(cfg.num_images, cfg.appearance_dim, generator=g)
    P["field.mlp_base.model.0.hash_table"] = table(cfg.main_grid)
    H, G = cfg.hidden_dim, cfg.geo_feat_dim
    enc = cfg.main_grid.num_levels * cfg.main_grid.features_per_level
    for name, dims in (
        ("field.mlp_base.model.1", [enc, H, 1 + G]),
        ("field.mlp_pred_normals", [12 + G, 64, 64, 64]),
        ("field.mlp_head", [16 + G + cfg.appearance_dim, 64, 64, 3]),
    ):
        for i in range(len(dims) - 1):
            w, b = _linear_init(g, dims[i + 1], dims[i])
            P[f"{name}.layers.{i}.weight"], P[f"{name}.layers.{i}.bias"] = w, b
    P["field.field_head_pred_normals.net.weight"], P["field.field_head_pred_normals.net.bias"] = _linear_init(g, 3, 64)
    for k, gc in enumerate(cfg.prop_grids):
        P[f"proposal_networks.{k}.encoding.hash_table"] = table(gc)
        enc_k = gc.num_levels * gc.features_per_level
        w, b = _linear_init(g, cfg.prop_hidden_dim, enc_k)
        P[f"proposal_networks.{k}.mlp_base.1.layers.0.weight"], P[f"proposal_networks.{k}.mlp_base.1.layers.0.bias"] = w, b
        w, b = _linear_init(g, 1, cfg.prop_hidden_dim)
        P[f"proposal_networks.{k}.mlp_base.1.layers.1.weight"], P[f"proposal_networks.{k}.mlp_base.1.layers.1.bias"] = w, b
    return P


def _layers(P: Dict[str, Tensor], prefix: str) -> Tuple[List[Tensor], List[Tensor]]:
    ws, bs, i = [], [], 0
    while f"{prefix}.layers.{i}.weight" in P:
        ws.append(P[f"{prefix}.layers.{i}.weight"])
        bs.append(P[f"{prefix}.layers.{i}.bias"])
        i += 1
    return ws, bs


# ----------------------------------------------------------------------------------------
# a8  fields
# ----------------------------------------------------------------------------------------


def sample_positions(origins: Tensor, directions: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """Frustum centre o + d*(s+e)/2 (NS/cameras/rays.py:55). origins/directions [B,3], starts/ends [B,S] -> [B,S,3]."""
    return origins[:, None, :] + directions[:, None, :] * (starts + ends)[..., None] / 2


def proposal_density(P: Dict[str, Tensor], k: int, gc: GridCfg, positions: Tensor) -> Tensor:
    """HashMLPDensityField.get_density (NS/fields/density_fields.py:93-116): positions [...,3] -> density [...]."""
    x, sel = normalized_positions(positions)
    feat = hash_encode(x.reshape(-1, 3), P[f"proposal_networks.{k}.encoding.hash_table"], level_scalings(gc), gc.log2_hashmap_size)
    ws, bs = _layers(P, f"proposal_networks.{k}.mlp_base.1")
    raw = mlp_forward(feat, ws, bs).reshape(sel.shape)
    return trunc_exp(raw) * sel


def nerfacto_field(
    P: Dict[str, Tensor],
    cfg: ModelCfg,
    origins: Tensor,
    directions: Tensor,
    starts: Tensor,
    ends: Tensor,
    camera_indices: Tensor,
    training: bool = True,
) -> Dict[str, Tensor]:
    """NerfactoField.forward(compute_normals=True) (NS/fields/base_field.py:114-133,
    nerfacto_field.py:199-297).  Returns density [B,S], rgb [B,S,3], normals [B,S,3], pred_normals [B,S,3]."""
    B, S = starts.shape
    gc = cfg.main_grid
    pos = sample_positions(origins, directions, starts, ends)
    with torch.enable_grad():
        x, sel = normalized_positions(pos)
        if not x.requires_grad:  # nerfacto_field.py:210-212: a leaf for the normals unless the rays already carry a graph (camera optimizer on:
            x = x.detach().requires_grad_(True)  # then the sample locations stay differentiable w.r.t. the pose deltas)
        feat = hash_encode(x.reshape(-1, 3), P["field.mlp_base.model.0.hash_table"], level_scalings(gc), gc.log2_hashmap_size)
        ws, bs = _layers(P, "field.mlp_base.model.1")
        h = mlp_forward(feat, ws, bs).reshape(B, S, -1)
        raw, geo = h[..., 0], h[..., 1:]
        density = trunc_exp(raw) * sel
        # base_field.py:92-99: normals = -normalize(d raw / d x), first order only
        (g,) = torch.autograd.grad(raw, x, grad_outputs=torch.ones_like(raw), retain_graph=True)
    normals = -F.normalize(g, dim=-1)

    d01 = (directions + 1.0) / 2.0  # base_field.py:142
    sh = sh_deg4(d01)[:, None, :].expand(B, S, 16)
    emb_w = P["field.embedding_appearance.embedding.weight"]
    if training:
        app = emb_w[camera_indices.reshape(-1).long()][:, None, :].expand(B, S, -1)
    else:  # use_average_appearance_embedding=True (nerfacto.py:112, nerfacto_field.py:243-246)
        app = (torch.ones(B, S, emb_w.shape[1]) * emb_w.mean(dim=0))
    out = {"density": density, "normals": normals, "selector": sel, "positions": pos, "x_normalized": x.detach()}
    if cfg.predict_normals:
        pe = posenc_2freq(pos)
        ws, bs = _layers(P, "field.mlp_pred_normals")
        hp = mlp_forward(torch.cat([pe, geo], dim=-1), ws, bs)
        pn = torch.tanh(hp @ P["field.field_head_pred_normals.net.weight"].t() + P["field.field_head_pred_normals.net.bias"])
        out["pred_normals"] = F.normalize(pn, dim=-1)  # field_heads.py:197-204
    ws, bs = _layers(P, "field.mlp_head")
    out["rgb"] = mlp_forward(torch.cat([sh, geo, app], dim=-1), ws, bs, "sigmoid")
    return out


# ----------------------------------------------------------------------------------------
# a9 / a10 / a12  samplers and weights
# ----------------------------------------------------------------------------------------


def spacing_fn(t: Tensor) -> Tensor:
    """UniformLinDispPiecewise s(t) (NS/model_components/ray_samplers.py:244)."""
    return torch.where(t < 1, t / 2, 1 - 1 / (2 * t))


def spacing_fn_inv(s: Tensor) -> Tensor:
    """t(s) (ray_samplers.py:245)."""
    return torch.where(s < 0.5, 2 * s, 1 / (2 - 2 * s))


def spacing_to_euclidean(bins: Tensor, nears: Tensor, fars: Tensor) -> Tensor:
    """ray_samplers.py:112-117; bins [B,n], nears/fars [B,1]."""
    s_near, s_far = spacing_fn(nears), spacing_fn(fars)
    return spacing_fn_inv(bins * s_far + (1 - bins) * s_near)


def uniform_spacing_bins(num_rays: int, num_samples: int, jitter: Optional[Tensor]) -> Tensor:
    """Initial spacing bins [B, S+1]; jitter [B,1] in [0,1) for the training single-jitter path, None in eval
    (ray_samplers.py:100-110)."""
    bins = torch.linspace(0.0, 1.0, num_samples + 1)[None, :]
    if jitter is not None:
        centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
        upper = torch.cat([centers, bins[..., -1:]], -1)
        lower = torch.cat([bins[..., :1], centers], -1)
        bins = lower + (upper - lower) * jitter
    else:
        bins = bins.expand(num_rays, -1)
    return bins


def get_weights(deltas: Tensor, densities: Tensor) -> Tensor:
    """alpha-compositing weights [B,S] (NS/cameras/rays.py:138-148)."""
    dd = deltas * densities
    alphas = 1 - torch.exp(-dd)
    trans = torch.cumsum(dd[..., :-1], dim=-1)
    trans = torch.cat([torch.zeros_like(trans[..., :1]), trans], dim=-1)
    return torch.nan_to_num(alphas * torch.exp(-trans))


def pdf_resample(
    weights: Tensor, existing_bins: Tensor, num_samples: int, jitter: Optional[Tensor], histogram_padding: float = 0.01, eps: float = 1e-5
) -> Dict[str, Tensor]:
    """PDFSampler (ray_samplers.py:305-362).  weights [B,S_in] (already annealed), existing_bins [B,S_in+1]
    spacing-space edges, jitter [B,1] (train, single jitter) or None (eval).
    Returns new spacing bins [B,num_samples+1] plus the integer intermediates (inds/below/above) and cdf/u."""
    num_bins = num_samples + 1
    w = weights + histogram_padding
    w_sum = torch.sum(w, dim=-1, keepdim=True)
    padding = torch.relu(eps - w_sum)
    w = w + padding / w.shape[-1]
    w_sum = w_sum + padding
    pdf = w / w_sum
    cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
    u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins)
    if jitter is not None:
        u = u.expand(*cdf.shape[:-1], num_bins) + jitter / num_bins
    else:
        u = (u + 1.0 / (2 * num_bins)).expand(*cdf.shape[:-1], num_bins)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, side="right")
    below = torch.clamp(inds - 1, 0, existing_bins.shape[-1] - 1)
    above = torch.clamp(inds, 0, existing_bins.shape[-1] - 1)
    cdf_g0, bins_g0 = torch.gather(cdf, -1, below), torch.gather(existing_bins, -1, below)
    cdf_g1, bins_g1 = torch.gather(cdf, -1, above), torch.gather(existing_bins, -1, above)
    t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
    bins = (bins_g0 + t * (bins_g1 - bins_g0)).detach()
    return {"bins": bins, "inds": inds, "below": below, "above": above, "cdf": cdf, "u": u}


def anneal_value(step: int, slope: float = 10.0, max_iters: int = 1000) -> float:
    """Proposal-weight anneal exponent (NS/models/nerfacto.py:258-270)."""
    x = float(np.clip(step / max_iters, 0, 1))
    return slope * x / ((slope - 1) * x + 1)


# ----------------------------------------------------------------------------------------
# a13-a15 renderers
# ----------------------------------------------------------------------------------------


def render_rgb(rgb: Tensor, weights: Tensor, training: bool = True) -> Tensor:
    """background 'last_sample' (NS/model_components/renderers.py:102-116,223-229). rgb [B,S,3], weights [B,S]."""
    if not training:
        rgb = torch.nan_to_num(rgb)
    comp = torch.sum(weights[..., None] * rgb, dim=-2)
    acc = torch.sum(weights, dim=-1, keepdim=True)
    comp = comp + rgb[..., -1, :] * (1.0 - acc)
    if not training:
        comp = comp.clamp(0.0, 1.0)
    return comp


def render_accumulation(weights: Tensor) -> Tensor:
    """renderers.py:314."""
    return torch.sum(weights, dim=-1, keepdim=True)


def render_depth_median(weights: Tensor, starts: Tensor, ends: Tensor) -> Tuple[Tensor, Tensor]:
    """renderers.py:353-362 -> (depth [B,1], median_index [B,1] int64)."""
    steps = (starts + ends) / 2
    cw = torch.cumsum(weights, dim=-1)
    split = torch.ones((*weights.shape[:-1], 1)) * 0.5
    idx = torch.searchsorted(cw, split, side="left")
    idx = torch.clamp(idx, 0, steps.shape[-1] - 1)
    return torch.gather(steps, dim=-1, index=idx), idx


def render_depth_expected(weights: Tensor, starts: Tensor, ends: Tensor) -> Tensor:
    """renderers.py:364-379 (clip bounds are min/max over the WHOLE batch)."""
    steps = (starts + ends) / 2
    depth = torch.sum(weights * steps, dim=-1, keepdim=True) / (torch.sum(weights, -1, keepdim=True) + 1e-10)
    return torch.clip(depth, steps.min(), steps.max())


def render_normals(normals: Tensor, weights: Tensor) -> Tensor:
    """NormalsRenderer + safe_normalize + NormalsShader (renderers.py:444-446, NS/utils/math.py:280-294,
    NS/model_components/shaders.py:74). Returns the (n+1)/2 colour-coded map [B,3]."""
    n = torch.sum(weights[..., None] * normals, dim=-2)
    n = n / (torch.norm(n, dim=-1, keepdim=True) + 1e-10)
    return (n + 1) / 2


# ----------------------------------------------------------------------------------------
# a16 losses (NS/model_components/losses.py)
# ----------------------------------------------------------------------------------------

EPS = 1.0e-7  # losses.py:37


def outer_bound(t0_starts, t0_ends, t1_starts, t1_ends, y1) -> Tuple[Tensor, Tensor, Tensor]:
    """losses.py:52-79 -> (y0_outer, idx_lo, idx_hi)."""
    cy1 = torch.cat([torch.zeros_like(y1[..., :1]), torch.cumsum(y1, dim=-1)], dim=-1)
    idx_lo = torch.searchsorted(t1_starts.contiguous(), t0_starts.contiguous(), side="right") - 1
    idx_lo = torch.clamp(idx_lo, min=0, max=y1.shape[-1] - 1)
    idx_hi = torch.searchsorted(t1_ends.contiguous(), t0_ends.contiguous(), side="right")
    idx_hi = torch.clamp(idx_hi, min=0, max=y1.shape[-1] - 1)
    cy1_lo = torch.take_along_dim(cy1[..., :-1], idx_lo, dim=-1)
    cy1_hi = torch.take_along_dim(cy1[..., 1:], idx_hi, dim=-1)
    return cy1_hi - cy1_lo, idx_lo, idx_hi


def interlevel_loss(weights_list: Sequence[Tensor], sdist_list: Sequence[Tensor]) -> Tensor:
    """losses.py:93-130; weights [B,S_k], sdist [B,S_k+1] per level; last entry = final level."""
    c = sdist_list[-1].detach()
    w = weights_list[-1].detach()
    loss = 0.0
    for cp, wp in zip(sdist_list[:-1], weights_list[:-1]):
        w_outer, _, _ = outer_bound(c[..., :-1], c[..., 1:], cp[..., :-1], cp[..., 1:], wp)
        loss = loss + torch.mean(torch.clip(w - w_outer, min=0) ** 2 / (w + EPS))
    return loss


def distortion_loss(weights: Tensor, sdist: Tensor) -> Tensor:
    """losses.py:134-153 on the final level."""
    ut = (sdist[..., 1:] + sdist[..., :-1]) / 2
    dut = torch.abs(ut[..., :, None] - ut[..., None, :])
    inter = torch.sum(weights * torch.sum(weights[..., None, :] * dut, dim=-1), dim=-1)
    intra = torch.sum(weights**2 * (sdist[..., 1:] - sdist[..., :-1]), dim=-1) / 3
    return torch.mean(inter + intra)


def ds_nerf_depth_loss(weights: Tensor, starts: Tensor, ends: Tensor, termination_depth: Tensor, directions_norm: Tensor, sigma: float) -> Tensor:
    """losses.py:224-246,313-319 with is_euclidean=False: depth [B,1] * directions_norm [B,1]; note 2*sigma, not 2*sigma^2."""
    D = termination_depth * directions_norm
    steps = (starts + ends) / 2
    lengths = ends - starts
    sig = torch.tensor([sigma])
    loss = -torch.log(weights + EPS) * torch.exp(-((steps - D) ** 2) / (2 * sig)) * lengths
    loss = loss.sum(-1, keepdim=True) * (D > 0)
    return torch.mean(loss)


def monosdf_normal_loss(normal_pred: Tensor, normal_gt: Tensor) -> Tensor:
    """losses.py:327-342."""
    g = F.normalize(normal_gt, p=2, dim=-1)
    p = F.normalize(normal_pred, p=2, dim=-1)
    return torch.abs(p - g).sum(dim=-1).mean() + (1.0 - torch.sum(p * g, dim=-1)).mean()


def orientation_loss(weights: Tensor, normals: Tensor, viewdirs: Tensor) -> Tensor:
    """losses.py:200-211 -> [B]."""
    n_dot_v = (normals * (-viewdirs)[..., None, :]).sum(dim=-1)
    return (weights * torch.fmin(torch.zeros_like(n_dot_v), n_dot_v) ** 2).sum(dim=-1)


def pred_normal_loss(weights: Tensor, normals: Tensor, pred_normals: Tensor) -> Tensor:
    """losses.py:214-221 -> [B]."""
    return (weights * (1.0 - torch.sum(normals * pred_normals, dim=-1))).sum(dim=-1)


# ----------------------------------------------------------------------------------------
# a11 + model: one mapping step (NS/models/nerfacto.py:288-381, depth_nerfacto.py:79-125,
# nerf_vo/mapping/nerfstudio_utils.py:333-350)
# ----------------------------------------------------------------------------------------


def mapping_forward(
    P: Dict[str, Tensor],
    cfg: ModelCfg,
    origins: Tensor,
    directions: Tensor,
    camera_indices: Tensor,
    jitters: Optional[Sequence[Tensor]] = None,
    anneal: float = 1.0,
    training: bool = True,
    prop_requires_grad: bool = True,
) -> Dict[str, object]:
    """NerfactoModel.get_outputs.  jitters = one [B,1] tensor per sampling level (3) in training; None in eval.
    Returns rendered maps plus the per-level lists the losses need."""
    B = origins.shape[0]
    near = cfg.near if training else 0.0  # NS/model_components/scene_colliders.py:188
    nears = torch.ones(B, 1) * near
    fars = torch.ones(B, 1) * cfg.far
    n_levels = len(cfg.num_proposal_samples)
    weights_list, sdist_list, starts_list, ends_list, aux = [], [], [], [], []
    weights = None
    sbins = None
    for lvl in range(n_levels + 1):
        is_prop = lvl < n_levels
        S = cfg.num_proposal_samples[lvl] if is_prop else cfg.num_nerf_samples
        j = jitters[lvl] if (training and jitters is not None) else None
        if lvl == 0:
            sbins = uniform_spacing_bins(B, S, j)
            if sbins.shape[0] != B:
                sbins = sbins.expand(B, -1)
        else:
            res = pdf_resample(torch.pow(weights, anneal), sbins, S, j, cfg.histogram_padding)  # ray_samplers.py:602
            sbins = res["bins"]
            aux.append(res)
        ebins = spacing_to_euclidean(sbins, nears, fars)
        starts, ends = ebins[..., :-1], ebins[..., 1:]
        if is_prop:
            pos = sample_positions(origins, directions, starts, ends)
            if prop_requires_grad:
                density = proposal_density(P, lvl, cfg.prop_grids[lvl], pos)
            else:
                with torch.no_grad():
                    density = proposal_density(P, lvl, cfg.prop_grids[lvl], pos)
            weights = get_weights(ends - starts, density)
            weights_list.append(weights)
            sdist_list.append(sbins)
            starts_list.append(starts)
            ends_list.append(ends)
    fo = nerfacto_field(P, cfg, origins, directions, starts, ends, camera_indices, training)
    weights = get_weights(ends - starts, fo["density"])
    weights_list.append(weights)
    sdist_list.append(sbins)
    starts_list.append(starts)
    ends_list.append(ends)
    out: Dict[str, object] = {
        "rgb": render_rgb(fo["rgb"], weights, training),
        "accumulation": render_accumulation(weights),
        "expected_depth": render_depth_expected(weights, starts, ends),
        "normals": render_normals(fo["normals"], weights),
        "weights_list": weights_list,
        "sdist_list": sdist_list,
        "starts_list": starts_list,
        "ends_list": ends_list,
        "field": fo,
        "pdf_aux": aux,
    }
    with torch.no_grad():
        out["depth"], out["depth_index"] = render_depth_median(weights, starts, ends)
        for i in range(n_levels):
            out[f"prop_depth_{i}"], _ = render_depth_median(weights_list[i], starts_list[i], ends_list[i])
    if cfg.predict_normals:
        out["pred_normals"] = render_normals(fo["pred_normals"], weights)
        if training:
            out["rendered_orientation_loss"] = orientation_loss(weights.detach(), fo["normals"], directions)
            out["rendered_pred_normal_loss"] = pred_normal_loss(weights.detach(), fo["normals"].detach(), fo["pred_normals"])
    return out


def mapping_losses(
    cfg: ModelCfg,
    out: Dict[str, object],
    rgb_gt: Tensor,
    depth_gt: Optional[Tensor],
    directions_norm: Optional[Tensor],
    normal_gt: Optional[Tensor] = None,
) -> Dict[str, Tensor]:
    """get_metrics_dict + get_loss_dict of ExtendedNerfactoModel(DepthNerfactoModel(NerfactoModel))."""
    L: Dict[str, Tensor] = {}
    L["rgb_loss"] = F.mse_loss(rgb_gt, out["rgb"])  # nerfacto.py:362
    wl, sl = out["weights_list"], out["sdist_list"]
    L["interlevel_loss"] = cfg.interlevel_loss_mult * interlevel_loss(wl, sl)
    L["distortion_loss"] = cfg.distortion_loss_mult * distortion_loss(wl[-1], sl[-1])
    if cfg.predict_normals:
        L["orientation_loss"] = cfg.orientation_loss_mult * torch.mean(out["rendered_orientation_loss"])
        L["pred_normal_loss"] = cfg.pred_normal_loss_mult * torch.mean(out["rendered_pred_normal_loss"])
    if depth_gt is not None:
        d = 0.0
        n = len(wl)
        for i in range(n):  # depth_nerfacto.py:93-103: all three weight sets
            d = d + ds_nerf_depth_loss(wl[i], out["starts_list"][i], out["ends_list"][i], depth_gt, directions_norm, cfg.depth_sigma) / n
        L["depth_loss"] = cfg.depth_loss_mult * d
    if normal_gt is not None and cfg.normal_loss_mult > 0.0:
        L["normal_loss"] = cfg.normal_loss_mult * monosdf_normal_loss(out["normals"], normal_gt)
    return L


def mapping_step(P, cfg, rays: Dict[str, Tensor], targets: Dict[str, Tensor], jitters, anneal=1.0, prop_requires_grad=True):
    """forward + losses + backward. P tensors must have requires_grad=True where gradients are wanted.
    Returns (outputs, loss_dict, total_loss); gradients land in P[k].grad."""
    out = mapping_forward(P, cfg, rays["origins"], rays["directions"], rays["camera_indices"], jitters, anneal, True, prop_requires_grad)
    L = mapping_losses(cfg, out, targets["rgb"], targets.get("depth"), rays.get("directions_norm"), targets.get("normal"))
    total = sum(L.values())
    total.backward()
    return out, L, total


# ----------------------------------------------------------------------------------------
# synthetic Replica-shaped inputs — shared by tests and bench
# ----------------------------------------------------------------------------------------


def synthetic_rays(num_rays: int, num_images: int = 192, seed: int = 1234, H: int = 680, W: int = 1200, fx: float = 600.0, fy: float = 600.0,
                   cx: float = 599.5, cy: float = 339.5) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    """Pinhole rays from `num_images` random cameras (origins U[-0.5,0.5]^3, uniform random rotations), pixels
    floor(rand*[K,H,W]) (NS/data/pixel_samplers.py:103-106); directions per NS/cameras/cameras.py:620-654,875-878."""
    g = torch.Generator().manual_seed(seed)
    cam_o = torch.rand(num_images, 3, generator=g) - 0.5
    q = F.normalize(torch.randn(num_images, 4, generator=g), dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1).reshape(-1, 3, 3)
    pix = torch.floor(torch.rand(num_rays, 3, generator=g) * torch.tensor([num_images, H, W])).long()
    c, py, px = pix.unbind(-1)
    dcam = torch.stack([(px + 0.5 - cx) / fx, -(py + 0.5 - cy) / fy, -torch.ones(num_rays)], dim=-1)
    dworld = torch.einsum("nij,nj->ni", R[c], dcam)
    dnorm = torch.linalg.norm(dworld, dim=-1, keepdim=True)
    rays = {
        "origins": cam_o[c].contiguous(),
        "directions": (dworld / dnorm).contiguous(),
        "camera_indices": c[:, None].contiguous(),
        "directions_norm": dnorm.contiguous(),
        "pixel_area": torch.full((num_rays, 1), 1.0 / (fx * fy)),
    }
    depth = torch.rand(num_rays, 1, generator=g) * 4.7 + 0.3
    depth = depth * (torch.rand(num_rays, 1, generator=g) > 0.1)
    targets = {
        "rgb": torch.rand(num_rays, 3, generator=g),
        "depth": depth,
        "normal": F.normalize(torch.randn(num_rays, 3, generator=g), dim=-1),
    }
    return rays, targets


def synthetic_jitters(num_rays: int, n_levels: int = 3, seed: int = 99) -> List[Tensor]:
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(num_rays, 1, generator=g) for _ in range(n_levels)]


# ----------------------------------------------------------------------------------------
# step prologue: pixel sampling, pixel gather, ray generation, camera-pose correction
# ----------------------------------------------------------------------------------------


def pixel_indices(u: Tensor, num_images: int, height: int, width: int) -> Tensor:
    """(rand(B,3) * [K,H,W]).long()  (NS/data/pixel_samplers.py:103-106): fp32 product, truncation toward zero."""
    return (u * torch.tensor([num_images, height, width], dtype=torch.float32)).long()


def generate_rays(indices: Tensor, intrinsics: Tensor, c2w: Tensor) -> Dict[str, Tensor]:
    """Perspective branch of Cameras.generate_rays (NS/cameras/cameras.py:596-633 image coordinates, :654 OpenCV->OpenGL
    flip, :780-785 directions, :865-892 rotation, normalisation, pixel area) on pixel centres `index + 0.5`
    (RayGenerator, NS/model_components/ray_generators.py:51-54; Cameras.get_image_coords :307-311).
    indices: [B,3] int64 (camera, row, col); intrinsics: [K,4] fx fy cx cy; c2w: [K,3|4,4]."""
    c, yi, xi = indices.unbind(-1)
    y, x = yi.float() + 0.5, xi.float() + 0.5
    fx, fy, cx, cy = intrinsics[c].unbind(-1)
    cs = torch.stack([torch.stack([(x - cx) / fx, (y - cy) / fy], -1), torch.stack([(x - cx + 1) / fx, (y - cy) / fy], -1),
                      torch.stack([(x - cx) / fx, (y - cy + 1) / fy], -1)], dim=0)  # [3,B,2]
    d = torch.stack([cs[..., 0], -cs[..., 1], -torch.ones_like(cs[..., 0])], dim=-1)  # [3,B,3]
    R = c2w[c][:, :3, :3]
    d = torch.sum(d[..., None, :] * R, dim=-1)
    norm = torch.maximum(torch.linalg.vector_norm(d, dim=-1, keepdim=True), torch.tensor([1e-8]))  # camera_utils.py:286-298, _EPS
    d = d / norm
    dx = torch.sqrt(torch.sum((d[0] - d[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((d[0] - d[2]) ** 2, dim=-1))
    return {"origins": c2w[c][:, :3, 3], "directions": d[0], "pixel_area": (dx * dy)[..., None], "camera_indices": c[:, None],
            "directions_norm": norm[0]}


def exp_map_so3xr3(t: Tensor) -> Tensor:
    """[R|t] of SO(3) x R^3 (NS/cameras/lie_groups.py:25-60): Rodrigues with the squared angle clamped at 1e-4."""
    w = t[:, 3:]
    ang = torch.clamp((w * w).sum(1), 1e-4).sqrt()
    f1 = ang.sin() / ang
    f2 = (1.0 - ang.cos()) / (ang * ang)
    z = torch.zeros_like(w[:, 0])
    K = torch.stack([z, -w[:, 2], w[:, 1], w[:, 2], z, -w[:, 0], -w[:, 1], w[:, 0], z], -1).reshape(-1, 3, 3)
    Rm = f1[:, None, None] * K + f2[:, None, None] * torch.bmm(K, K) + torch.eye(3)[None]
    return torch.cat([Rm, t[:, :3, None]], dim=-1)


def exp_map_se3(t: Tensor) -> Tensor:
    """se(3) -> SE(3), tangent ordered (translation, rotation) like lietorch.SE3.exp (NS/cameras/lie_groups.py:63-120:
    the reference calls lietorch, absent here — PARITY UNPINNED for this mode; the closed form below is the one the
    reference keeps in comments :70-117 and is cross-checked against torch.linalg.matrix_exp in the tests)."""
    v, w = t[:, :3].double(), t[:, 3:].double()
    th2 = (w * w).sum(1)
    small = th2 < 1e-8
    ths = torch.where(small, torch.ones_like(th2), th2).sqrt()  # sqrt only where it is differentiable
    A = torch.where(small, 1 - th2 / 6, ths.sin() / ths)
    Bc = torch.where(small, 0.5 - th2 / 24, (1 - ths.cos()) / (ths * ths))
    C = torch.where(small, 1.0 / 6 - th2 / 120, (ths - ths.sin()) / (ths * ths * ths))
    z = torch.zeros_like(th2)
    K = torch.stack([z, -w[:, 2], w[:, 1], w[:, 2], z, -w[:, 0], -w[:, 1], w[:, 0], z], -1).reshape(-1, 3, 3)
    K2 = torch.bmm(K, K)
    eye = torch.eye(3, dtype=torch.float64)[None]
    Rm = eye + A[:, None, None] * K + Bc[:, None, None] * K2
    V = eye + Bc[:, None, None] * K + C[:, None, None] * K2
    return torch.cat([Rm, torch.bmm(V, v[:, :, None])], dim=-1).float()


def apply_pose_correction(origins: Tensor, directions: Tensor, camera_indices: Tensor, pose_adjustment: Optional[Tensor], mode: str):
    """CameraOptimizer.apply_to_raybundle (NS/cameras/camera_optimizers.py:138-147): origins + t, R @ directions."""
    if mode == "off" or pose_adjustment is None:
        return origins, directions
    M = (exp_map_so3xr3 if mode == "SO3xR3" else exp_map_se3)(pose_adjustment[camera_indices.reshape(-1)])
    return origins + M[:, :3, 3], torch.bmm(M[:, :3, :3], directions[..., None]).squeeze(-1)


def next_train_batch(u: Tensor, intrinsics: Tensor, extrinsics: Tensor, frames_color: Tensor, frames_depth: Tensor,
                     frames_normal: Optional[Tensor] = None, pose_adjustment: Optional[Tensor] = None, mode: str = "off"):
    """DynamicDataManager.next_train (nerf_vo/mapping/nerfstudio_utils.py:295-300) + the pose correction NerfactoModel.get_outputs
    applies in training (NS/models/nerfacto.py:290-291): dataset dict (:133-155, normals (R^-1 n + 1)/2 — evaluated at the sampled
    pixels only, the per-pixel solve is independent of the rest of the frame), PixelSampler.collate_image_dataset_batch
    (NS/data/pixel_samplers.py:170-219), RayGenerator.  Returns (rays, batch)."""
    K, H, W, _ = frames_color.shape
    idx = pixel_indices(u, K, H, W)
    c, y, x = idx.unbind(-1)
    batch = {"indices": idx, "image": frames_color[c, y, x], "depth_image": frames_depth[c, y, x]}
    if frames_normal is not None:
        n = torch.linalg.solve(extrinsics[c, :3, :3], frames_normal[c, y, x][..., None]).squeeze(-1)
        batch["normal_image"] = (n + 1) / 2
    rays = generate_rays(idx, intrinsics, extrinsics[:, :3])
    rays["origins"], rays["directions"] = apply_pose_correction(rays["origins"], rays["directions"], rays["camera_indices"], pose_adjustment, mode)
    return rays, batch


# ----------------------------------------------------------------------------------------
# evaluation frame render
# ----------------------------------------------------------------------------------------


def frame_rays(intrinsics: Tensor, extrinsics_std: np.ndarray, height: int, width: int) -> Dict[str, Tensor]:
    """Full-frame bundle of NerfstudioRenderer.render_frame (evaluation/nerf_renderer.py:134-158): standard -> NeRF axis convention
    (columns 1:3 negated), then Cameras.generate_rays(camera_indices=0, keep_shape=True) on every pixel centre, row-major."""
    ext = np.array(extrinsics_std, dtype=np.float64, copy=True)
    ext[0:3, 1:3] *= -1
    c2w = torch.tensor(ext, dtype=torch.float32)[None, :3]
    ys, xs = torch.meshgrid(torch.arange(height), torch.arange(width), indexing="ij")
    idx = torch.stack([torch.zeros_like(ys), ys, xs], dim=-1).reshape(-1, 3)
    return generate_rays(idx, intrinsics.reshape(1, 4), c2w)


def finalize_frame(rgb: Tensor, depth: Tensor, directions_norm: Tensor) -> Tuple[np.ndarray, np.ndarray]:
    """(rgb * 255).astype(uint8) and depth / directions_norm (evaluation/nerf_renderer.py:160-167; DepthNerfacto branch)."""
    color = (rgb.numpy() * 255).astype(np.uint8)
    return color, (depth / directions_norm).numpy()[..., 0]


def depth_scale_sums(depth_gt: np.ndarray, depth_pred: np.ndarray) -> Tuple[float, float, int]:
    """Masked sums behind depth_gt[mask].mean() / depth_pred[mask].mean() (evaluation/renderer.py:88-93)."""
    mask = (depth_gt > 0) * (depth_pred > 0) * (depth_gt < 5) * (depth_pred < 5)
    return float(depth_gt[mask].astype(np.float64).sum()), float(depth_pred[mask].astype(np.float64).sum()), int(mask.sum())


def depth_to_uint16(depth: np.ndarray, scale_pred2gt: float, depth_scale: float) -> np.ndarray:
    """(raw_depth * scale_pred2gt * depth_scale).astype(uint16) in fp32 (evaluation/renderer.py:116-119)."""
    return (depth.astype(np.float32) * np.float32(scale_pred2gt) * np.float32(depth_scale)).astype(np.uint16)


def render_frame(P: Dict[str, Tensor], cfg: ModelCfg, intrinsics: Tensor, extrinsics_std: np.ndarray, height: int, width: int, chunk: int = 4096):
    """NerfstudioRenderer.render_frame: eval-mode forward in `chunk`-ray slices (NS/models/base_model.py:164-192) -> (uint8 colour, depth)."""
    rays = frame_rays(intrinsics, extrinsics_std, height, width)
    n = height * width
    rgb, depth = [], []
    with torch.no_grad():
        for i in range(0, n, chunk):
            o = mapping_forward(P, cfg, rays["origins"][i:i + chunk], rays["directions"][i:i + chunk], rays["camera_indices"][i:i + chunk], None, 1.0, training=False)
            rgb.append(o["rgb"])
            depth.append(o["depth"])
    color, d = finalize_frame(torch.cat(rgb), torch.cat(depth), rays["directions_norm"])
    return color.reshape(height, width, 3), d.reshape(height, width)


def point_cloud_batch(origins: Tensor, directions: Tensor, depth: Tensor, accumulation: Tensor, normals_coded: Optional[Tensor] = None,
                      bounding_box_min=None, bounding_box_max=None, reorient_normals: bool = False):
    """One iteration of generate_point_cloud's loop plus its final re-orientation (NS/exporter/exporter_utils.py:130-180, 222-226), the way
    NerfstudioRenderer.render_mesh drives it (evaluation/nerf_renderer.py:188-203): returns the COMPACTED (points, rgb-mask indices, normals)."""
    point = origins + directions * depth
    view = directions
    mask = accumulation[..., 0] > 0.5 if accumulation.dim() == 2 else accumulation > 0.5  # get_rgba_image: alpha = accumulation
    idx = torch.nonzero(mask)[:, 0]
    point, view = point[mask], view[mask]
    normal = None
    if normals_coded is not None:
        normal = (normals_coded * 2.0) - 1.0
        normal = normal[mask]
    if bounding_box_min is not None:
        comp_l, comp_m = torch.tensor(bounding_box_min), torch.tensor(bounding_box_max)
        assert torch.all(comp_l < comp_m)
        m2 = torch.all(torch.concat([point > comp_l, point < comp_m], dim=-1), dim=-1)
        point, view, idx = point[m2], view[m2], idx[m2]
        if normal is not None:
            normal = normal[m2]
    if reorient_normals and normal is not None:
        normal = normal.clone()
        flip = torch.sum(view * normal, dim=-1) > 0
        normal[flip] *= -1
    return point, idx, normal
