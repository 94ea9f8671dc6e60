"""Functional operators + autograd glue over the C-ABI kernels.  torch is used for memory, streams and the
autograd graph only; every arithmetic step of the hot path runs in libnvo_b200.so.  No CPU fallback."""
from __future__ import annotations

import os

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import call, check

_env_cache: dict = {}


def env_flag(name: str, default: bool) -> bool:
    """Experiment switch read from the environment ONCE per process, not on every step."""
    if name not in _env_cache:
        v = os.environ.get(name)
        _env_cache[name] = default if v is None else v == "1"
    return _env_cache[name]


# ------------------------------------------------------------------------------------------------
# side streams for gradient-leaf kernels
# ------------------------------------------------------------------------------------------------


class _LeafStreams:
    """Kernels that only scatter into the trainer's flat gradient buffer (hash-table scatter, fused proposal backward) have no
    consumer inside the backward pass, so they may run on side streams next to the remaining (latency-bound) backward chain.
    The trainer enables this, and joins the streams before the optimizer / all-reduce.  Tensors such a kernel reads are kept
    alive until the join, so the caching allocator cannot hand their blocks to main-stream work that would race with it."""

    def __init__(self):
        self.enabled = False
        self.streams: List[torch.cuda.Stream] = []
        self.used: List[torch.cuda.Stream] = []
        self.refs: list = []
        self.next = 0
        # set by the trainer for the duration of one backward pass:
        #   after_field_backward(): called on the side stream that carries the main hash-table scatter, right after its launch;
        #   defer_event: the fused proposal backward waits for it (keeps those long kernels out of the field's backward chain)
        self.after_field_backward = None
        self.defer_event = None

    def enable(self, n: int = 3, n_levels: int = 2):
        self.streams = [torch.cuda.Stream() for _ in range(n)]
        # one stream per proposal level: the level's forward (density + weights) is issued there, so autograd runs the level's
        # BACKWARD there too, ordered only after the gradients it consumes — i.e. concurrently with the main field's backward
        if os.environ.get("NVO_ONE_LEVEL_STREAM", "0") == "1":
            one = torch.cuda.Stream(priority=int(os.environ.get("NVO_LEVEL_PRIORITY", "0")))
            self.level_streams = [one for _ in range(n_levels)]
        else:
            self.level_streams = [torch.cuda.Stream(priority=int(os.environ.get("NVO_LEVEL_PRIORITY", "0"))) for _ in range(n_levels)]
        # outputs nothing inside the step consumes (proposal depth maps): their own streams, so they never queue in front of a loss kernel
        self.aux_streams = [torch.cuda.Stream() for _ in range(2)]
        self.next_aux = 0
        # independent branches of the field's forward (density-gradient normals, predicted-normals network) next to the colour head:
        # each is a chain of latency-bound launches that leaves most of the SMs idle
        self.branch_streams = [torch.cuda.Stream(priority=-1) for _ in range(2)]
        # the main hash-table scatter is the tail of the step's critical path (the optimizer waits for it): its own high-priority stream, so
        # its CTAs are placed ahead of the proposal networks' backward and scatters that are pending at the same time
        self.critical_stream = torch.cuda.Stream(priority=-1) if os.environ.get("NVO_SCATTER_PRIORITY", "1") == "1" else None
        self.enabled = True

    def on_level_stream(self) -> bool:
        cur = torch.cuda.current_stream()
        return any(cur == st for st in getattr(self, "level_streams", []))

    def fork(self, *keepalive, aux: bool = False, critical: bool = False):
        """Returns a context manager running its body on the next side stream, ordered after the current stream."""
        cur = torch.cuda.current_stream()
        if critical and getattr(self, "critical_stream", None) is not None:
            st = self.critical_stream
        elif aux:
            st = self.aux_streams[self.next_aux % len(self.aux_streams)]
            self.next_aux += 1
        else:
            st = self.streams[self.next % len(self.streams)]
            self.next += 1
        st.wait_stream(cur)
        if st not in self.used:
            self.used.append(st)
        self.refs.append(keepalive)
        return torch.cuda.stream(st)

    def join(self):
        cur = torch.cuda.current_stream()
        for st in self.used:
            cur.wait_stream(st)
        self.used.clear()
        self.refs.clear()
        self.next = 0
        self.next_aux = 0


leaf_streams = _LeafStreams()


# ------------------------------------------------------------------------------------------------
# hash grid
# ------------------------------------------------------------------------------------------------


def torch_level_scalings(num_levels: int, min_res: int, max_res: int) -> torch.Tensor:
    """The fp32 per-level scale exactly as the reference evaluates it (float32 pow then floor;
    NS/field_components/encodings.py:346-349). Never recomputed on the device."""
    levels = torch.arange(num_levels)
    growth = np.exp((np.log(max_res) - np.log(min_res)) / (num_levels - 1)) if num_levels > 1 else 1
    return torch.floor(min_res * growth**levels)


def growth_level_scalings(num_levels: int, base_resolution: int, per_level_scale: float) -> torch.Tensor:
    """Same, from the tcnn JSON keys (base_resolution, per_level_scale)."""
    levels = torch.arange(num_levels)
    return torch.floor(base_resolution * np.float64(per_level_scale) ** levels)


@dataclass
class GridSpec:
    n_levels: int
    log2_T: int
    scalings: Tuple[float, ...]
    features_per_level: int = 2

    def __post_init__(self):
        if self.features_per_level != 2:
            raise RuntimeError("nvo_b200 hash grid supports n_features_per_level == 2 only")
        self._desc = {}

    @property
    def n_rows(self) -> int:
        return self.n_levels << self.log2_T

    @property
    def out_dim(self) -> int:
        return self.n_levels * 2

    def desc(self, table_dtype, out_dtype):
        k = (table_dtype, out_dtype)
        if k not in self._desc:
            self._desc[k] = _lib.make_grid_desc(self.n_levels, self.log2_T, self.scalings, table_dtype, out_dtype)
        return self._desc[k]


def _check_grid_inputs(x, table, spec: GridSpec):
    check(x, "grid input", torch.float32, (None, 3))
    if table.dtype not in (torch.float32, torch.float16):
        raise RuntimeError(f"hash table must be float32 or float16, got {table.dtype}")
    check(table, "hash table", table.dtype)
    if table.numel() != spec.n_rows * 2:
        raise RuntimeError(f"hash table has {table.numel()} elements, expected {spec.n_rows * 2}")
    if x.device != table.device:
        raise RuntimeError("grid input and hash table must be on the same device")


def tmh_numel(n: int, kpad: int) -> int:
    """fp16 elements of a TMH buffer ("tile-major half", the tensor-core MLP's operand layout): ceil(n/128) tiles of 128 x kpad."""
    return (n + 127) // 128 * 128 * kpad


def grid_forward(x, table, spec: GridSpec, out_dtype=torch.float32):
    """out_dtype: torch.float32 / torch.float16 -> [n, 2L] row-major; "tmh" -> fp16 TMH tiles (columns padded to 16, zero filled)."""
    _check_grid_inputs(x, table, spec)
    if out_dtype == "tmh":
        y = torch.empty(tmh_numel(x.shape[0], (spec.out_dim + 15) // 16 * 16), dtype=torch.float16, device=x.device)
    else:
        y = torch.empty((x.shape[0], spec.out_dim), dtype=out_dtype, device=x.device)
    call("nvo_grid_forward", spec.desc(table.dtype, out_dtype), x.shape[0], x, table, y)
    return y


def grid_backward(x, dy, spec: GridSpec, dtable=None, n_rows=None, tmf: bool = False):
    """Scatter dy into a fp32 gradient table (allocated zero-filled when not given).  tmf: dy is the fp32 tile-major buffer
    the tensor-core MLP backward writes ([tile][2L][128]) instead of [n, 2L] row-major."""
    check(x, "grid input", torch.float32, (None, 3))
    if tmf:
        check(dy, "grid dy (tmf)", torch.float32, (tmh_numel(x.shape[0], spec.out_dim),))
    else:
        check(dy, "grid dy", dy.dtype, (x.shape[0], spec.out_dim))
    if dtable is None:
        dtable = torch.zeros((spec.n_rows, 2), dtype=torch.float32, device=x.device)
    call("nvo_grid_backward", spec.desc(torch.float32, "tmf" if tmf else dy.dtype), x.shape[0], x, dy, dtable)
    return dtable


def grid_backward_input(x, table, dy, spec: GridSpec, tmf: bool = False):
    _check_grid_inputs(x, table, spec)
    if tmf:
        check(dy, "grid dy (tmf)", torch.float32, (tmh_numel(x.shape[0], spec.out_dim),))
    else:
        check(dy, "grid dy", dy.dtype, (x.shape[0], spec.out_dim))
    dx = torch.empty_like(x)
    call("nvo_grid_backward_input", spec.desc(table.dtype, "tmf" if tmf else dy.dtype), x.shape[0], x, table, dy, dx)
    return dx


def grid_forward_jac(x, table, spec: GridSpec):
    """TMH features plus the saved d(feature)/d(x) (fp16, 3 x the feature buffer; include/nvo_b200.h: nvo_grid_forward_jac)."""
    _check_grid_inputs(x, table, spec)
    numel = tmh_numel(x.shape[0], (spec.out_dim + 15) // 16 * 16)
    y = torch.empty(numel, dtype=torch.float16, device=x.device)
    jac = torch.empty(3 * numel, dtype=torch.float16, device=x.device)
    call("nvo_grid_forward_jac", spec.desc(table.dtype, "tmh"), x.shape[0], x, table, y, jac)
    return y, jac


def grid_jac_dx(jac, dy, spec: GridSpec, n: int, normalize_scale: float = 0.0, eps: float = 1e-12):
    """dx [n,3] from the saved derivatives and the tile-major fp32 dy; normalize_scale != 0 fuses scale * v / max(|v|, eps)."""
    check(jac, "grid jacobian", torch.float16, (3 * tmh_numel(n, (spec.out_dim + 15) // 16 * 16),))
    check(dy, "grid dy (tmf)", torch.float32, (tmh_numel(n, spec.out_dim),))
    dx = torch.empty((n, 3), dtype=torch.float32, device=dy.device)
    call("nvo_grid_jac_dx", spec.desc(torch.float32, "tmf"), n, jac, dy, normalize_scale, eps, dx)
    return dx


def grid_indices(x, spec: GridSpec):
    check(x, "grid input", torch.float32, (None, 3))
    idx = torch.empty((x.shape[0], spec.n_levels, 8), dtype=torch.int64, device=x.device)
    call("nvo_grid_indices", spec.desc(torch.float32, torch.float32), x.shape[0], x, idx)
    return idx


class _GridEncode(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, table, spec, out_dtype):
        x = x.contiguous()
        y = grid_forward(x, table, spec, out_dtype)
        ctx.save_for_backward(x, table)
        ctx.spec = spec
        # fused gradient accumulation: a trainer may attach a preallocated fp32 gradient buffer to the parameter
        ctx.main_grad = getattr(table, "_nvo_main_grad", None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, table = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dtable = None
        if ctx.needs_input_grad[1] and ctx.main_grad is not None:
            grid_backward(x, dy, ctx.spec, dtable=ctx.main_grad)  # scatter straight into the trainer's flat gradient
        elif ctx.needs_input_grad[1]:
            dtable = grid_backward(x, dy, ctx.spec).view(table.shape)
            if table.dtype != torch.float32:
                dtable = dtable.to(table.dtype)
        if ctx.needs_input_grad[0]:
            dx = grid_backward_input(x, table, dy, ctx.spec)
        return dx, dtable, None, None


def grid_encode(x, table, spec: GridSpec, out_dtype=torch.float32):
    return _GridEncode.apply(x, table, spec, out_dtype)


# ------------------------------------------------------------------------------------------------
# MLP
# ------------------------------------------------------------------------------------------------


@dataclass
class MlpSpec:
    in_dim: int
    dims: Tuple[int, ...]
    activation: str = "relu"
    out_activation: str = "none"
    acts: Optional[Tuple[str, ...]] = None  # per-layer override

    def __post_init__(self):
        if self.acts is None:
            self.acts = tuple([self.activation] * (len(self.dims) - 1) + [self.out_activation])
        self.desc = _lib.make_mlp_desc(self.in_dim, self.dims, self.acts)
        ins = [self.in_dim] + list(self.dims[:-1])
        self.shapes: List[Tuple[Tuple[int, int], Tuple[int]]] = [((o, i), (o,)) for i, o in zip(ins, self.dims)]
        self.n_params = sum(o * i + o for i, o in zip(ins, self.dims))
        self.saved_per_sample = sum(self.dims[:-1])

    @property
    def out_dim(self) -> int:
        return self.dims[-1]

    def offsets(self):
        off, res = 0, []
        for (w, b) in self.shapes:
            res.append((off, off + w[0] * w[1]))
            off += w[0] * w[1]
            res.append((off, off + b[0]))
            off += b[0]
        return res


def flat_alias(params: Sequence[torch.Tensor]) -> Optional[torch.Tensor]:
    """If `params` are back-to-back views of one storage (our modules allocate them that way), return the flat
    fp32 tensor aliasing them; otherwise None."""
    p0 = params[0]
    expect = p0.data_ptr()
    for p in params:
        if p.dtype != torch.float32 or not p.is_contiguous() or p.data_ptr() != expect or p.untyped_storage().data_ptr() != p0.untyped_storage().data_ptr():
            return None
        expect += p.numel() * 4
    n = sum(p.numel() for p in params)
    return torch.empty(0, dtype=torch.float32, device=p0.device).set_(p0.untyped_storage(), p0.storage_offset(), (n,), (1,))


def mlp_forward(x, flat, spec: MlpSpec, save: bool, row_mask=None):
    check(x, "mlp input", torch.float32, (None, spec.in_dim))
    check(flat, "mlp params", torch.float32, (spec.n_params,))
    n = x.shape[0]
    if row_mask is not None:
        check(row_mask, "row_mask", torch.float32, (n,))
    y = torch.empty((n, spec.out_dim), dtype=torch.float32, device=x.device)
    saved = torch.empty((n, spec.saved_per_sample), dtype=torch.float32, device=x.device) if (save and spec.saved_per_sample) else None
    call("nvo_mlp_forward", spec.desc, n, x, flat, row_mask, y, saved)
    return y, saved


def mlp_backward(x, flat, saved, y, dy, spec: MlpSpec, need_dx: bool, need_dparams: bool, dflat=None, row_mask=None):
    n = x.shape[0]
    check(dy, "mlp dy", torch.float32, (n, spec.out_dim))
    dx = torch.empty_like(x) if need_dx else None
    if need_dparams and dflat is None:
        dflat = torch.zeros(spec.n_params, dtype=torch.float32, device=x.device)
    call("nvo_mlp_backward", spec.desc, n, x, flat, saved, y, row_mask, dy, dx, dflat if need_dparams else None)
    return dx, dflat


class _MlpApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, spec, row_mask, *params):
        x = x.contiguous()
        with torch.no_grad():
            flat = flat_alias(params)
            if flat is None:
                flat = torch.cat([p.reshape(-1).float() for p in params])
        need = x.requires_grad or any(p.requires_grad for p in params)
        y, saved = mlp_forward(x, flat, spec, save=need, row_mask=row_mask)
        ctx.save_for_backward(x, flat, saved, y, row_mask)
        ctx.spec = spec
        ctx.n_tensors = len(params)
        ctx.main_grad = getattr(params[0], "_nvo_main_grad", None)  # flat fp32 gradient of all `params`, or None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, flat, saved, y, row_mask = ctx.saved_tensors
        spec = ctx.spec
        need_dx = ctx.needs_input_grad[0]
        need_dp = any(ctx.needs_input_grad[3:])
        dx, dflat = mlp_backward(x, flat, saved, y, dy.contiguous(), spec, need_dx, need_dp, dflat=ctx.main_grad, row_mask=row_mask)
        grads = [None] * ctx.n_tensors
        if need_dp and ctx.main_grad is None:
            grads = []
            for (a, b), (ws, bs) in zip(_pairs(spec.offsets()), spec.shapes):
                grads.append(dflat[a[0]:a[1]].view(ws))
                grads.append(dflat[b[0]:b[1]].view(bs))
        return (dx, None, None, *grads)


def _pairs(seq):
    it = iter(seq)
    return list(zip(it, it))


def mlp_apply(x, spec: MlpSpec, params: Sequence[torch.Tensor], row_mask=None):
    """params = [W0, b0, W1, b1, ...] in torch layout; row_mask [n] of 0/1 multiplies the output rows."""
    return _MlpApply.apply(x, spec, row_mask, *params)


# ------------------------------------------------------------------------------------------------
# fused proposal density field (grid + 16-wide MLP + trunc_exp * selector in one kernel each way)
# ------------------------------------------------------------------------------------------------


def prop_density_supported(gspec: GridSpec, mspec: MlpSpec) -> bool:
    if len(mspec.dims) != 2 or mspec.dims[1] != 1 or mspec.acts[0] != "relu" or mspec.acts[1] != "trunc_exp" or mspec.in_dim != gspec.out_dim:
        return False
    return bool(_lib.load().nvo_prop_density_supported(gspec.n_levels, mspec.dims[0], 2))


# constant-memory banks of the fused proposal field that hold their network for the step in flight: slot -> (event recorded behind the
# upload, data_ptr of the parameters uploaded).  Set by prop_density_preload (the trainer, at the start of a step), cleared by clear_prepacked().
_prop_resident: dict = {}


def prop_density_preload(slot: int, params) -> None:
    """Uploads the proposal network `params` into constant-memory bank `slot` on a side stream (next to whatever the current stream does
    next); _PropDensity's forward / backward of that slot then wait for the upload's event instead of copying the parameters themselves."""
    flat = _flat_of(params)
    slot = int(slot) % 4
    if leaf_streams.enabled:
        with leaf_streams.fork(flat):
            call("nvo_prop_density_upload", slot, flat)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
    else:
        call("nvo_prop_density_upload", slot, flat)
        ev = None
    _prop_resident[slot] = (ev, flat.data_ptr())


def _prop_params_arg(slot: int, flat):
    """None (after ordering the current stream behind the upload) when bank `slot` already holds `flat`, else `flat`."""
    hit = _prop_resident.get(slot)
    if hit is None or hit[1] != flat.data_ptr():
        return flat
    if hit[0] is not None:
        torch.cuda.current_stream().wait_event(hit[0])
    return None


class _PropDensity(torch.autograd.Function):
    """density [B,S] at the samples of `iv` on rays (origins, directions), or at explicit `positions` [B*S,3]."""

    @staticmethod
    def forward(ctx, table, gspec, hidden, slot, origins, directions, iv, positions, B, S, *params):
        flat = _flat_of(params)
        dev = table.device
        need = any(ctx.needs_input_grad)  # False under no_grad (eval, frozen proposal steps): no features saved
        n = B * S
        density = torch.empty((B, S), dtype=torch.float32, device=dev)
        feat = torch.empty(int(_lib.load().nvo_prop_density_feat_floats(gspec.n_levels, n)), dtype=torch.float32, device=dev) if need else None
        if positions is not None:
            positions = check(positions.reshape(-1, 3).contiguous(), "positions", torch.float32, (n, 3))
            s = e = None
            stride = 0
        else:
            check(origins, "origins", torch.float32, (B, 3))
            check(directions, "directions", torch.float32, (B, 3))
            s, e, stride = iv.triple()
        call("nvo_prop_density_forward", gspec.desc(table.dtype, torch.float32), hidden, slot, B, S, origins, directions, s, e, stride, positions, table,
             _prop_params_arg(slot, flat), density, feat)
        ctx.save_for_backward(table, flat, feat, origins, directions, positions)
        ctx.iv, ctx.gspec, ctx.hidden, ctx.slot, ctx.B, ctx.S, ctx.n_tensors = iv, gspec, hidden, slot, B, S, len(params)
        ctx.table_main_grad = getattr(table, "_nvo_main_grad", None)
        ctx.mlp_main_grad = getattr(params[0], "_nvo_main_grad", None)
        ctx.mspec_shapes = [tuple(p.shape) for p in params]
        ctx.sink = ray_grad_sink if (positions is None and any(ctx.needs_input_grad)) else None
        return density

    @staticmethod
    def backward(ctx, ddensity):
        table, flat, feat, origins, directions, positions = ctx.saved_tensors
        need_dt, need_dp = ctx.needs_input_grad[0], any(ctx.needs_input_grad[10:])
        # camera-pose optimisation: d loss / d (origins, directions) through this level's sample positions
        need_rays = positions is None and (ctx.needs_input_grad[4] or ctx.needs_input_grad[5] or ctx.sink is not None)
        d_o = d_d = None
        dev = table.device
        dtable = dflat = None
        if need_dt:
            dtable = ctx.table_main_grad if ctx.table_main_grad is not None else torch.zeros(table.shape, dtype=torch.float32, device=dev)
        if need_dp:
            dflat = ctx.mlp_main_grad if ctx.mlp_main_grad is not None else torch.zeros(flat.numel(), dtype=torch.float32, device=dev)
        if positions is not None:
            s = e = None
            stride = 0
        else:
            s, e, stride = ctx.iv.triple()
        ddensity = ddensity.contiguous()
        params_arg = _prop_params_arg(ctx.slot, flat)
        args = ("nvo_prop_density_backward", ctx.gspec.desc(table.dtype, torch.float32), ctx.hidden, ctx.slot, ctx.B, ctx.S, origins, directions, s, e, stride, positions,
                params_arg, feat, ddensity, dtable, dflat)
        split = (need_dt and env_flag("NVO_PROP_BWD_SPLIT", True)) or need_rays
        if need_rays:
            d_o, d_d = ctx.sink if ctx.sink is not None else (torch.zeros((ctx.B, 3), dtype=torch.float32, device=dev),
                                                              torch.zeros((ctx.B, 3), dtype=torch.float32, device=dev))

        def run():
            if not split:
                call(*args)
                return
            # two lean kernels instead of the fused one: MLP part -> tile-major feature gradients + positions, then the long-run scatter
            n = ctx.B * ctx.S
            dft = torch.empty(tmh_numel(n, ctx.gspec.out_dim), dtype=torch.float32, device=dev)
            xq = torch.empty(((n + 127) // 128 * 128, 3), dtype=torch.float32, device=dev)
            call("nvo_prop_density_backward_split", ctx.gspec.desc(table.dtype, torch.float32), ctx.hidden, ctx.slot, ctx.B, ctx.S, params_arg, feat, ddensity, dflat,
                 dft, xq)
            def ray_part():
                dxn = grid_backward_input(xq[:n], table, dft, ctx.gspec, tmf=True)
                position_backward(origins, directions, ctx.iv, dxn, d_o, d_d)

            if need_rays and ctx.sink is not None and leaf_streams.enabled:
                # the gather pass for d loss / d positions (L2-read bound) next to the table scatter (reduction bound), both reading dft;
                # joined with the other leaf streams before the optimizer touches the table
                with leaf_streams.fork(xq, dft, table, origins, directions, ctx.iv):
                    ray_part()
            if need_dt:
                grid_backward(xq[:n], dft, ctx.gspec, dtable=dtable.view(-1, 2), tmf=True)
            if need_rays and not (ctx.sink is not None and leaf_streams.enabled):
                ray_part()

        if leaf_streams.defer_event is not None:
            torch.cuda.current_stream().wait_event(leaf_streams.defer_event)
        if (leaf_streams.enabled and not leaf_streams.on_level_stream() and (not need_dt or dtable is ctx.table_main_grad)
                and (not need_dp or dflat is ctx.mlp_main_grad) and (not need_rays or ctx.sink is not None)):
            with leaf_streams.fork(table, flat, feat, origins, directions, positions, ddensity, ctx.iv):
                run()
        else:
            run()
        if dtable is ctx.table_main_grad:
            dtable = None
        elif dtable is not None and table.dtype != torch.float32:
            dtable = dtable.to(table.dtype)
        grads = [None] * ctx.n_tensors
        if need_dp and ctx.mlp_main_grad is None:
            off = 0
            grads = []
            for shp in ctx.mspec_shapes:
                k = int(np.prod(shp))
                grads.append(dflat[off:off + k].view(shp))
                off += k
        if ctx.sink is not None:
            d_o = d_d = None
        return (dtable, None, None, None, d_o if ctx.needs_input_grad[4] else None, d_d if ctx.needs_input_grad[5] else None, None, None, None, None, *grads)


def prop_density(table, gspec: GridSpec, mspec: MlpSpec, params, B: int, S: int, origins=None, directions=None, iv=None, positions=None, slot: int = 0):
    """slot: constant-memory bank (0..3) for the MLP parameters; give networks that may run concurrently distinct slots."""
    return _PropDensity.apply(table, gspec, mspec.dims[0], int(slot) % 4, origins, directions, iv, positions, B, S, *params)


# ------------------------------------------------------------------------------------------------
# field element-wise operators
# ------------------------------------------------------------------------------------------------


def contract_normalize(positions):
    """SceneContraction(L-inf) + (x+2)/4 + selector masking: positions [...,3] -> (x [n,3], selector [n])."""
    p = check(positions.reshape(-1, 3).contiguous(), "positions", torch.float32)
    x = torch.empty_like(p)
    sel = torch.empty(p.shape[0], dtype=torch.float32, device=p.device)
    call("nvo_contract_forward", p.shape[0], p, x, sel)
    return x, sel


def sh4(directions):
    d = check(directions.reshape(-1, 3).contiguous(), "directions", torch.float32)
    out = torch.empty((d.shape[0], 16), dtype=torch.float32, device=d.device)
    call("nvo_sh4_forward", d.shape[0], d, out)
    return out


def frequency(x, n_freq: int):
    x = check(x.contiguous(), "frequency input", torch.float32)
    n, d = x.shape
    out = torch.empty((n, d * n_freq * 2), dtype=torch.float32, device=x.device)
    call("nvo_frequency_forward", n, d, n_freq, x, out)
    return out


class _TruncExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = check(x.float().contiguous(), "trunc_exp input", torch.float32)
        y = torch.empty_like(x)
        call("nvo_trunc_exp_forward", x.numel(), x, y)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        call("nvo_trunc_exp_backward", x.numel(), x, g.contiguous(), dx)
        return dx


trunc_exp = _TruncExp.apply


class _Normalize3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, scale, eps):
        shape = v.shape
        v2 = check(v.reshape(-1, 3).contiguous(), "normalize input", torch.float32)
        out = torch.empty_like(v2)
        call("nvo_normalize3_forward", v2.shape[0], v2, scale, eps, out)
        ctx.save_for_backward(v2)
        ctx.scale, ctx.eps, ctx.shape = scale, eps, shape
        return out.view(shape)

    @staticmethod
    def backward(ctx, g):
        (v2,) = ctx.saved_tensors
        dv = torch.empty_like(v2)
        call("nvo_normalize3_backward", v2.shape[0], v2, g.reshape(-1, 3).contiguous(), ctx.scale, ctx.eps, dv)
        return dv.view(ctx.shape), None, None


def normalize3(v, scale: float = 1.0, eps: float = 1e-12):
    return _Normalize3.apply(v, float(scale), float(eps))


class _FieldAssemble(torch.autograd.Function):
    """h [n,16], selector [n], directions [B,3], positions [n,3], cam_idx [B] int64 | None, embedding [K,32] | mean [32]
    -> density [n], head_in [n,63], pn_in [n,27] | None"""

    @staticmethod
    def forward(ctx, h, embedding, selector, directions, positions, cam_idx, B, S, want_pn):
        ctx.set_materialize_grads(False)
        n = B * S
        h = check(h.contiguous(), "mlp_base output", torch.float32, (n, 16))
        embedding = check(embedding.contiguous(), "appearance embedding", torch.float32)
        dev = h.device
        density = torch.empty(n, dtype=torch.float32, device=dev)
        head_in = torch.empty((n, 63), dtype=torch.float32, device=dev)
        pn_in = torch.empty((n, 27), dtype=torch.float32, device=dev) if want_pn else None
        call("nvo_field_assemble_forward", B, S, h, selector, directions, positions, cam_idx, embedding, 0, density, head_in, pn_in)
        ctx.save_for_backward(h, selector, cam_idx)
        ctx.B, ctx.S, ctx.emb_shape = B, S, embedding.shape
        ctx.main_grad = getattr(embedding, "_nvo_main_grad", None)
        return density, head_in, pn_in

    @staticmethod
    def backward(ctx, ddensity, dhead_in, dpn_in):
        h, selector, cam_idx = ctx.saved_tensors
        dev = h.device
        dh = torch.empty_like(h)
        if dhead_in is None:
            dhead_in = torch.zeros((h.shape[0], 63), dtype=torch.float32, device=dev)
        demb = None
        if ctx.needs_input_grad[1] and ctx.main_grad is not None and cam_idx is not None:
            demb = ctx.main_grad
        elif ctx.needs_input_grad[1]:
            demb = torch.zeros(ctx.emb_shape, dtype=torch.float32, device=dev)
            if cam_idx is None:
                # eval-style mean embedding: every sample contributes to the single vector
                demb = dhead_in[:, 31:].sum(0).reshape(ctx.emb_shape)
        c = lambda t: None if t is None else t.contiguous()
        call("nvo_field_assemble_backward", ctx.B, ctx.S, h, selector, cam_idx, c(ddensity), dhead_in.contiguous(), c(dpn_in), 0, dh,
             demb if cam_idx is not None else None, None)
        if demb is ctx.main_grad:
            demb = None
        return dh, demb, None, None, None, None, None, None, None


def field_assemble(h, embedding, selector, directions, positions, cam_idx, B: int, S: int, want_pn: bool):
    return _FieldAssemble.apply(h, embedding, selector, directions, positions, cam_idx, B, S, want_pn)


# ------------------------------------------------------------------------------------------------
# intervals helper: (starts, ends, stride) triple the per-ray kernels take
# ------------------------------------------------------------------------------------------------


class Intervals:
    """Euclidean sample intervals of B rays x S samples, either as compact edges ebins [B,S+1] or as separate
    starts/ends [B,S]."""

    def __init__(self, ebins: Optional[torch.Tensor] = None, starts: Optional[torch.Tensor] = None, ends: Optional[torch.Tensor] = None):
        if ebins is not None:
            self.ebins = check(ebins, "ebins", torch.float32)
            self.B, self.S = ebins.shape[0], ebins.shape[1] - 1
            self._starts = self._ends = None
        else:
            self.ebins = None
            self._starts = check(starts.contiguous(), "starts", torch.float32)
            self._ends = check(ends.contiguous(), "ends", torch.float32)
            self.B, self.S = self._starts.shape

    def triple(self):
        if self.ebins is not None:
            return self.ebins.data_ptr(), self.ebins.data_ptr() + 4, self.S + 1
        return self._starts.data_ptr(), self._ends.data_ptr(), self.S

    def keepalive(self):
        return (self.ebins, self._starts, self._ends)

    @property
    def starts(self):
        return self.ebins[:, :-1] if self.ebins is not None else self._starts

    @property
    def ends(self):
        return self.ebins[:, 1:] if self.ebins is not None else self._ends

    @property
    def device(self):
        return self.ebins.device if self.ebins is not None else self._starts.device


_const_cache = {}


def _cached_linspace(key, make, dev):
    """Small host-evaluated constant tables (torch.linspace on the CPU, exactly as the reference computes them), uploaded
    once per (shape, device) so that steady-state steps issue no H2D copies and stay CUDA-graph capturable."""
    k = (key, str(dev))
    if k not in _const_cache:
        _const_cache[k] = make().to(dev)
    return _const_cache[k]


def sample_uniform(num_samples: int, nears, fars, jitter=None):
    """-> (sdist [B,S+1], ebins [B,S+1])"""
    B = nears.shape[0]
    dev = nears.device
    nears = check(nears.reshape(-1).contiguous(), "nears", torch.float32, (B,))
    fars = check(fars.reshape(-1).contiguous(), "fars", torch.float32, (B,))
    base = _cached_linspace(("uniform", num_samples), lambda: torch.linspace(0.0, 1.0, num_samples + 1), dev)  # ray_samplers.py:100
    if jitter is not None:
        jitter = check(jitter.reshape(-1).contiguous(), "jitter", torch.float32, (B,))
    sdist = torch.empty((B, num_samples + 1), dtype=torch.float32, device=dev)
    ebins = torch.empty_like(sdist)
    call("nvo_sample_uniform", B, num_samples, base, jitter, nears, fars, sdist, ebins)
    return sdist, ebins


def sample_positions(origins, directions, iv: Intervals):
    check(origins, "origins", torch.float32, (iv.B, 3))
    check(directions, "directions", torch.float32, (iv.B, 3))
    pos = torch.empty((iv.B, iv.S, 3), dtype=torch.float32, device=origins.device)
    s, e, stride = iv.triple()
    call("nvo_sample_positions", iv.B, iv.S, origins, directions, s, e, stride, pos)
    return pos


class _Weights(torch.autograd.Function):
    @staticmethod
    def forward(ctx, density, iv: Intervals):
        density = check(density.contiguous(), "density", torch.float32, (iv.B, iv.S))
        w = torch.empty_like(density)
        s, e, stride = iv.triple()
        call("nvo_weights_forward", iv.B, iv.S, s, e, stride, density, w)
        ctx.save_for_backward(density)
        ctx.iv = iv
        return w

    @staticmethod
    def backward(ctx, dw):
        (density,) = ctx.saved_tensors
        iv = ctx.iv
        dd = torch.empty_like(density)
        s, e, stride = iv.triple()
        call("nvo_weights_backward", iv.B, iv.S, s, e, stride, density, dw.contiguous(), dd)
        return dd, None


def weights_from_density(density, iv: Intervals):
    """density [B,S] -> weights [B,S] (RaySamples.get_weights)."""
    return _Weights.apply(density, iv)


def pdf_resample(weights, sdist_in, num_samples: int, nears, fars, jitter=None, anneal=1.0, histogram_padding: float = 0.01,
                 return_inds: bool = False):
    """-> (sdist_out [B,S_out+1], ebins_out [B,S_out+1][, inds int32 [B,S_out+1]]); no gradient (the reference detaches).
    anneal: python float, or a device float32 [1] tensor the kernel reads (graph-replayable anneal schedule)."""
    B, S_in = weights.shape
    dev = weights.device
    weights = check(weights.detach().contiguous(), "weights", torch.float32, (B, S_in))
    sdist_in = check(sdist_in.contiguous(), "sdist", torch.float32, (B, S_in + 1))
    nears = check(nears.reshape(-1).contiguous(), "nears", torch.float32, (B,))
    fars = check(fars.reshape(-1).contiguous(), "fars", torch.float32, (B,))
    n = num_samples + 1

    def make_u():  # ray_samplers.py:317 / :327 evaluated by torch on the host
        u = torch.linspace(0.0, 1.0 - (1.0 / n), steps=n)
        return u + 1.0 / (2 * n) if jitter is None else u

    u = _cached_linspace(("pdf_u", n, jitter is None), make_u, dev)
    if jitter is not None:
        jitter = check(jitter.reshape(-1).contiguous(), "jitter", torch.float32, (B,))
    sdist = torch.empty((B, n), dtype=torch.float32, device=dev)
    ebins = torch.empty_like(sdist)
    inds = torch.empty((B, n), dtype=torch.int32, device=dev) if return_inds else None
    anneal_dev = None
    if isinstance(anneal, torch.Tensor):
        anneal_dev, anneal = check(anneal, "anneal", torch.float32, (1,)), 1.0
    call("nvo_pdf_resample", B, S_in, num_samples, weights, sdist_in, u, jitter, float(anneal), anneal_dev, float(histogram_padding), nears, fars, sdist,
         ebins, inds)
    return (sdist, ebins, inds) if return_inds else (sdist, ebins)


class _Render(torch.autograd.Function):
    """All renderers in one pass.  Outputs: rgb[B,3], acc[B,1], depth_expected[B,1], depth_median[B,1], median_idx[B,1] (int),
    normals[B,3], pred_normals[B,3] (None where the input is absent)."""

    @staticmethod
    def forward(ctx, weights, rgb, normals, pred_normals, iv: Intervals, eval_mode: bool, want_median: bool):
        ctx.set_materialize_grads(False)
        B, S = iv.B, iv.S
        dev = weights.device
        weights = check(weights.contiguous(), "weights", torch.float32, (B, S))
        f = lambda t, name: None if t is None else check(t.contiguous(), name, torch.float32, (B, S, 3))
        rgb, normals, pred_normals = f(rgb, "rgb"), f(normals, "normals"), f(pred_normals, "pred_normals")
        new = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        o_rgb = new(B, 3) if rgb is not None else None
        o_acc, o_dexp = new(B, 1), new(B, 1)
        o_dmed = new(B, 1) if want_median else None
        o_midx = torch.empty((B, 1), dtype=torch.int32, device=dev) if want_median else None
        o_n = new(B, 3) if normals is not None else None
        o_pn = new(B, 3) if pred_normals is not None else None
        minmax = torch.empty(2, dtype=torch.float32, device=dev)  # initialised on the device by nvo_render_forward
        s, e, stride = iv.triple()
        call("nvo_render_forward", B, S, int(eval_mode), s, e, stride, weights, rgb, normals, pred_normals, o_rgb, o_acc, o_dexp, minmax, o_dmed, o_midx,
             o_n, o_pn)
        call("nvo_clip_depth", B, minmax, o_dexp)
        ctx.save_for_backward(weights, rgb, normals, pred_normals, minmax)
        ctx.iv = iv
        outs = (o_rgb, o_acc, o_dexp, o_dmed, o_midx, o_n, o_pn)
        ctx.mark_non_differentiable(*[t for t in (o_dmed, o_midx) if t is not None])
        return outs

    @staticmethod
    def backward(ctx, d_rgb, d_acc, d_dexp, d_dmed, d_midx, d_n, d_pn):
        weights, rgb, normals, pred_normals, minmax = ctx.saved_tensors
        iv = ctx.iv
        B, S = iv.B, iv.S
        c = lambda t: None if t is None else t.contiguous()
        dw = torch.empty_like(weights)
        drgb = torch.empty_like(rgb) if (rgb is not None and ctx.needs_input_grad[1] and d_rgb is not None) else None
        # no upstream gradient on the rendered pred-normal map -> no gradient for the per-sample predictions (lets the
        # producer skip the whole mlp_pred_normals backward, as when pred_normal_loss_mult == 0)
        dpn = torch.empty_like(pred_normals) if (pred_normals is not None and ctx.needs_input_grad[3] and d_pn is not None) else None
        s, e, stride = iv.triple()
        call("nvo_render_backward", B, S, s, e, stride, weights, rgb, normals, pred_normals, c(d_rgb), c(d_acc), c(d_dexp), minmax, c(d_n), c(d_pn), 0, dw,
             drgb, dpn)
        return dw, drgb, None, dpn, None, None, None


def render(weights, iv: Intervals, rgb=None, normals=None, pred_normals=None, eval_mode=False, want_median=True):
    return _Render.apply(weights, rgb, normals, pred_normals, iv, eval_mode, want_median)


# ------------------------------------------------------------------------------------------------
# losses (scalar outputs are 0-dim tensors on the device; no host sync anywhere)
# ------------------------------------------------------------------------------------------------


def _scalar_like(t):
    return torch.zeros(1, dtype=torch.float32, device=t.device)


class _Distortion(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, sdist):
        B, S = weights.shape
        weights = check(weights.contiguous(), "weights", torch.float32, (B, S))
        sdist = check(sdist.contiguous(), "sdist", torch.float32, (B, S + 1))
        loss = _scalar_like(weights)
        call("nvo_distortion_loss_forward", B, S, weights, sdist, loss)
        ctx.save_for_backward(weights, sdist)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        weights, sdist = ctx.saved_tensors
        B, S = weights.shape
        dw = torch.zeros_like(weights)
        call("nvo_distortion_loss_backward", B, S, weights, sdist, g.reshape(1).float().contiguous(), 1.0, dw)
        return dw, None


class _Interlevel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, c, wp, cp):
        B, S = w.shape
        Sp = wp.shape[1]
        w = check(w.detach().contiguous(), "w", torch.float32, (B, S))
        c = check(c.detach().contiguous(), "c", torch.float32, (B, S + 1))
        wp = check(wp.contiguous(), "wp", torch.float32, (B, Sp))
        cp = check(cp.contiguous(), "cp", torch.float32, (B, Sp + 1))
        loss = _scalar_like(w)
        call("nvo_interlevel_loss_forward", B, S, Sp, w, c, wp, cp, loss, None, None)
        ctx.save_for_backward(w, c, wp, cp)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        w, c, wp, cp = ctx.saved_tensors
        B, S = w.shape
        Sp = wp.shape[1]
        dwp = torch.zeros_like(wp)
        call("nvo_interlevel_loss_backward", B, S, Sp, w, c, wp, cp, g.reshape(1).float().contiguous(), 1.0, dwp)
        return None, None, dwp, None


class _DepthLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, iv: Intervals, depth_gt, dnorm, sigma: float):
        B, S = iv.B, iv.S
        weights = check(weights.contiguous(), "weights", torch.float32, (B, S))
        depth_gt = check(depth_gt.reshape(-1).contiguous(), "termination_depth", torch.float32, (B,))
        dnorm = check(dnorm.reshape(-1).contiguous(), "directions_norm", torch.float32, (B,))
        loss = _scalar_like(weights)
        s, e, stride = iv.triple()
        call("nvo_depth_loss_forward", B, S, weights, s, e, stride, depth_gt, dnorm, float(sigma), loss)
        ctx.save_for_backward(weights, depth_gt, dnorm)
        ctx.iv, ctx.sigma = iv, float(sigma)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        weights, depth_gt, dnorm = ctx.saved_tensors
        iv = ctx.iv
        dw = torch.zeros_like(weights)
        s, e, stride = iv.triple()
        call("nvo_depth_loss_backward", iv.B, iv.S, weights, s, e, stride, depth_gt, dnorm, ctx.sigma, g.reshape(1).float().contiguous(), 1.0, dw)
        return dw, None, None, None, None


class _Mse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        pred = check(pred.contiguous(), "pred", torch.float32)
        target = check(target.contiguous(), "target", torch.float32, tuple(pred.shape))
        loss = _scalar_like(pred)
        d = torch.empty_like(pred)
        call("nvo_mse_loss", pred.numel(), pred, target, 1.0, loss, d)
        ctx.save_for_backward(d)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None


class _NormalLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt):
        B = pred.shape[0]
        pred = check(pred.contiguous(), "normal_pred", torch.float32, (B, 3))
        gt = check(gt.contiguous(), "normal_gt", torch.float32, (B, 3))
        loss = _scalar_like(pred)
        d = torch.empty_like(pred)
        call("nvo_normal_loss", B, pred, gt, 1.0, loss, d)
        ctx.save_for_backward(d)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None


def distortion_loss_op(weights, sdist):
    return _Distortion.apply(weights, sdist)


def interlevel_loss_op(w, c, wp, cp):
    return _Interlevel.apply(w, c, wp, cp)


def interlevel_indices(w, c, wp, cp):
    """Test hook: the clamped searchsorted indices (idx_lo, idx_hi) the interlevel loss uses."""
    B, S = w.shape
    Sp = wp.shape[1]
    lo = torch.empty((B, S), dtype=torch.int32, device=w.device)
    hi = torch.empty_like(lo)
    loss = _scalar_like(w)
    call("nvo_interlevel_loss_forward", B, S, Sp, w.contiguous(), c.contiguous(), wp.contiguous(), cp.contiguous(), loss, lo, hi)
    return lo, hi


def depth_loss_op(weights, iv: Intervals, depth_gt, directions_norm, sigma: float):
    return _DepthLoss.apply(weights, iv, depth_gt, directions_norm, sigma)


def mse_loss_op(pred, target):
    return _Mse.apply(pred, target)


def normal_loss_op(pred, gt):
    return _NormalLoss.apply(pred, gt)


def _run_levels(level) -> None:
    """level(0), level(1) on two helper streams, level(2) on the current one, joined before returning (plain sequential calls when
    side streams are not enabled).  Every tensor involved outlives the call, so no allocator hazards arise."""
    if not leaf_streams.enabled:
        for i in range(3):
            level(i)
        return
    main = torch.cuda.current_stream()
    helpers = leaf_streams.streams[:2]
    for i, st in enumerate(helpers):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            level(i)
    level(2)
    for st in helpers:
        main.wait_stream(st)


class _FusedStepLosses(torch.autograd.Function):
    """Every loss of the NeRF-VO mapping step (nerf_vo/mapping/nerfstudio.py:71-82: rgb MSE, interlevel, distortion, DS-NeRF depth
    over the three weight sets, MonoSDF normal) in ONE autograd node: the forward kernels add their batch means into a 5-vector,
    the backward kernels accumulate multiplier * dL/dweights straight into one zero-filled buffer per level.  Same kernels and
    arithmetic as the individual loss ops; what disappears is the ~40 scalar torch kernels that glue them together
    (multiplier products, zero fills, gradient accumulation adds, their backward twins).
    terms = [rgb, interlevel, distortion, depth (sum over levels), normal] (unweighted); total = sum(mults * terms)."""

    @staticmethod
    def forward(ctx, w0, w1, w2, rgb, normals_img, spec):
        ctx.set_materialize_grads(False)
        ws = [check(w.reshape(w.shape[0], -1).contiguous(), "weights", torch.float32) for w in (w0, w1, w2)]
        B = ws[0].shape[0]
        dev = ws[0].device
        rgb = check(rgb.contiguous(), "rgb", torch.float32, (B, 3))
        d_rgb = torch.empty_like(rgb)
        d_n = None
        if spec["normal_gt"] is not None and normals_img is not None:
            normals_img = check(normals_img.contiguous(), "normals", torch.float32, (B, 3))
            d_n = torch.empty_like(normals_img)
        wf, cf = ws[2], spec["sdist"][2]

        # eager_grads (the trainer's promise that `total` is differentiated with grad_output == 1, i.e. total.backward()): the gradient
        # kernels are launched right behind their forward twins, on the same per-level streams, instead of as a second wave after the
        # autograd engine has turned around — they only read the weights, so nothing orders them behind the scalar total.
        eager = bool(spec.get("eager_grads")) and any(ctx.needs_input_grad[:3])
        dws = None
        if eager:
            sizes = [w.shape[1] for w in ws]
            # one launch for the whole wave (csrc/losses.cu: k_step_losses): terms, weighted total and every gradient
            flat = torch.empty(B * sum(sizes), dtype=torch.float32, device=dev)
            dws, off = [], 0
            for n_s in sizes:
                dws.append(flat[off:off + B * n_s].view(B, n_s))
                off += B * n_s
            state = torch.zeros(8, dtype=torch.float32, device=dev)  # terms[5], total, ticket (uint32 zero), pad
            m = spec["mults"]
            use_d = spec["depth_gt"] is not None
            ivs = [spec["iv"][i].triple() if use_d else (None, None, 0) for i in range(3)]
            call("nvo_step_losses", B, sizes[0], sizes[1], sizes[2], ws[0], ws[1], ws[2], spec["sdist"][0], spec["sdist"][1], spec["sdist"][2],
                 ivs[0][0], ivs[0][1], ivs[0][2], ivs[1][0], ivs[1][1], ivs[1][2], ivs[2][0], ivs[2][1], ivs[2][2], rgb, spec["rgb_gt"],
                 normals_img if d_n is not None else None, spec["normal_gt"] if d_n is not None else None, spec["depth_gt"], spec["dnorm"], spec["sigma"],
                 m[0], m[1], m[2], m[3] if use_d else 0.0, m[4] if d_n is not None else 0.0, state, state[5:], state[6:], dws[0], dws[1], dws[2], d_rgb, d_n)
            terms, total = state[:5], state[5]
            ctx.save_for_backward(dws[0], dws[1], dws[2], d_rgb, d_n)
            ctx.eager = True
            ctx.spec = spec
            ctx.mark_non_differentiable(terms)
            return total, terms
        terms = torch.zeros(5, dtype=torch.float32, device=dev)

        def level(i):  # everything that reads weight set i: the kernels are tiny (4096 warps), so the three levels run side by side
            if i < 2:
                call("nvo_interlevel_loss_forward", B, wf.shape[1], ws[i].shape[1], wf, cf, ws[i], spec["sdist"][i], terms[1:], None, None)
            else:
                call("nvo_mse_loss", rgb.numel(), rgb, spec["rgb_gt"], spec["mults"][0], terms, d_rgb)
                call("nvo_distortion_loss_forward", B, wf.shape[1], wf, cf, terms[2:])
                if d_n is not None:
                    call("nvo_normal_loss", B, normals_img, spec["normal_gt"], spec["mults"][4], terms[4:], d_n)
            if spec["depth_gt"] is not None:
                s, e, stride = spec["iv"][i].triple()
                call("nvo_depth_loss_forward", B, ws[i].shape[1], ws[i], s, e, stride, spec["depth_gt"], spec["dnorm"], spec["sigma"], terms[3:])

        _run_levels(level)
        total = torch.dot(terms, spec["mults_dev"])
        ctx.save_for_backward(*ws, d_rgb, d_n)
        ctx.eager = False
        ctx.spec = spec
        ctx.mark_non_differentiable(terms)
        return total, terms

    @staticmethod
    def backward(ctx, g, _g_terms):
        w0, w1, w2, d_rgb, d_n = ctx.saved_tensors
        ws, spec = [w0, w1, w2], ctx.spec
        if ctx.eager:  # gradients were produced next to the forward, for grad_output == 1
            shp = spec["w_shapes"]
            return (w0.view(shp[0]), w1.view(shp[1]), w2.view(shp[2]), d_rgb, d_n, None)
        B = w0.shape[0]
        sizes = [w.shape[1] for w in ws]
        g = g.reshape(1).float().contiguous()
        flat = torch.zeros(B * sum(sizes), dtype=torch.float32, device=w0.device)
        dws, off = [], 0
        for n_s in sizes:
            dws.append(flat[off:off + B * n_s].view(B, n_s))
            off += B * n_s
        m = spec["mults"]
        wf, cf = ws[2], spec["sdist"][2]

        def level(i):  # each level owns its gradient buffer: no two streams touch the same one
            if i < 2:
                call("nvo_interlevel_loss_backward", B, wf.shape[1], sizes[i], wf, cf, ws[i], spec["sdist"][i], g, m[1], dws[i])
            else:
                call("nvo_distortion_loss_backward", B, wf.shape[1], wf, cf, g, m[2], dws[2])
            if spec["depth_gt"] is not None:
                s, e, stride = spec["iv"][i].triple()
                call("nvo_depth_loss_backward", B, sizes[i], ws[i], s, e, stride, spec["depth_gt"], spec["dnorm"], spec["sigma"], g, m[3], dws[i])

        _run_levels(level)
        shp = spec["w_shapes"]
        return (dws[0].view(shp[0]), dws[1].view(shp[1]), dws[2].view(shp[2]), d_rgb * g, None if d_n is None else d_n * g, None)


def fused_step_losses(weights_list, sdist_list, iv_list, rgb, rgb_gt, normals_img=None, normal_gt=None, depth_gt=None, directions_norm=None,
                      sigma: float = 0.001, mults=(1.0, 1.0, 0.002, 0.001 / 3, 5e-6), eager_grads: bool = False):
    """mults = multipliers of [rgb, interlevel, distortion, depth (already divided by the number of levels), normal].
    eager_grads=True: the caller will call total.backward() (grad_output exactly 1): gradients are computed alongside the forward."""
    dev = rgb.device
    f = lambda t, shape: None if t is None else check(t.reshape(shape).contiguous(), "target", torch.float32)
    B = rgb.shape[0]
    key = ("loss_mults", tuple(float(x) for x in mults))
    mults_dev = _cached_linspace(key, lambda: torch.tensor([float(x) for x in mults], dtype=torch.float32), dev)
    spec = {"sdist": [check(s.contiguous(), "sdist", torch.float32) for s in sdist_list], "iv": list(iv_list), "rgb_gt": f(rgb_gt, (B, 3)),
            "normal_gt": f(normal_gt, (B, 3)), "depth_gt": f(depth_gt, (B,)), "dnorm": f(directions_norm, (B,)), "sigma": float(sigma),
            "mults": [float(x) for x in mults], "mults_dev": mults_dev, "w_shapes": [tuple(w.shape) for w in weights_list], "eager_grads": bool(eager_grads)}
    return _FusedStepLosses.apply(weights_list[0], weights_list[1], weights_list[2], rgb, normals_img, spec)


# ------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr: float, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, grad_scale: float = 1.0):
    """Fused dense Adam over flat fp32 buffers; `step` is a device int32 [1] tensor, incremented by the call."""
    n = params.numel()
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        check(t, name, torch.float32, (n,))
    check(step, "step", torch.int32, (1,))
    call("nvo_adam_step", n, params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale)


def adam_step_decay(params, grads, exp_avg, exp_avg_sq, step, lr_init: float, lr_final: float, max_steps: int, beta1: float = 0.9, beta2: float = 0.999,
                    eps: float = 1e-8, grad_scale: float = 1.0):
    """adam_step under ExponentialDecayScheduler (NS/engine/schedulers.py:109-141, no warm-up): the learning rate is evaluated on the device
    from `step`, so a captured CUDA graph follows the schedule."""
    n = params.numel()
    for t, name in ((params, "params"), (grads, "grads"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        check(t, name, torch.float32, (n,))
    check(step, "step", torch.int32, (1,))
    call("nvo_adam_step_decay", n, params, grads, exp_avg, exp_avg_sq, step, lr_init, lr_final, int(max_steps), beta1, beta2, eps, grad_scale)


def pose_regularizer(pose_adjustment, trans_l2_penalty: float, rot_l2_penalty: float, loss=None, d_pose=None, scale: float = 1.0):
    """CameraOptimizer.get_loss_dict (camera_optimizers.py:149-155): adds the regulariser to `loss` (device scalar) and scale * its gradient
    to d_pose [K,6]."""
    check(pose_adjustment, "pose_adjustment", torch.float32, (None, 6))
    call("nvo_pose_regularizer", pose_adjustment.shape[0], pose_adjustment, trans_l2_penalty, rot_l2_penalty, scale, loss, d_pose)


# ------------------------------------------------------------------------------------------------
# tensor-core (tcgen05) MLP path: fp16 operands, fp32 accumulate, for widths <= 64 and depth <= 4
# ------------------------------------------------------------------------------------------------


def tc_eligible(spec: MlpSpec) -> bool:
    return len(spec.dims) <= 4 and spec.in_dim <= 64 and all(d <= 64 for d in spec.dims)


def tc_in_pad(spec: MlpSpec) -> int:
    return (spec.in_dim + 15) // 16 * 16


def cast_pad_f16(x, spec: MlpSpec):
    """fp32 [n,in_dim] row-major -> fp16 TMH tiles (in_pad columns, whole tiles, zero padded)."""
    check(x, "mlp input", torch.float32, (None, spec.in_dim))
    out = torch.empty(tmh_numel(x.shape[0], tc_in_pad(spec)), dtype=torch.float16, device=x.device)
    call("nvo_cast_pad_f16", x.shape[0], spec.in_dim, tc_in_pad(spec), x, out)
    return out


def _tc_saved_bytes(spec: MlpSpec, n: int) -> int:
    import ctypes

    return int(_lib.load().nvo_mlp_tc_saved_bytes(ctypes.addressof(spec.desc), n))


def tc_pack_weights(flat, spec: MlpSpec):
    """fp32 torch-layout parameters -> the packed fp16 weight image the tensor-core kernels bulk-copy into shared memory."""
    import ctypes

    check(flat, "mlp params", torch.float32, (spec.n_params,))
    hit = _prepacked.get((flat.data_ptr(), id(spec)))
    if hit is not None:
        torch.cuda.current_stream().wait_event(hit[1])  # packed ahead on a side stream (prepack_weights)
        return hit[0]
    nbytes = getattr(spec, "_wimage_bytes", None)
    if nbytes is None:
        nbytes = spec._wimage_bytes = int(_lib.load().nvo_mlp_tc_wimage_bytes(ctypes.addressof(spec.desc)))
    img = torch.empty(nbytes, dtype=torch.uint8, device=flat.device)
    call("nvo_mlp_tc_pack_weights", spec.desc, flat, img)
    return img


# weight images packed ahead of their use: (parameter pointer, spec) -> (image, ready event).  The trainer fills it at the start of a
# step on a side stream (the parameters only change in the optimizer), so the three repack launches leave the forward's critical chain;
# it is cleared before the optimizer runs.
_prepacked: dict = {}


def prepack_weights(nets) -> None:
    """nets: iterable of (params list, MlpSpec).  Images are allocated on the CURRENT stream (their consumers' stream) and written on a
    side stream; tc_pack_weights() hands them out after waiting for the ready event."""
    import ctypes

    jobs = []
    for params, spec in nets:
        flat = _flat_of(params)
        nbytes = getattr(spec, "_wimage_bytes", None)
        if nbytes is None:
            nbytes = spec._wimage_bytes = int(_lib.load().nvo_mlp_tc_wimage_bytes(ctypes.addressof(spec.desc)))
        jobs.append((flat, spec, torch.empty(nbytes, dtype=torch.uint8, device=flat.device)))
    if not jobs:
        return
    with leaf_streams.fork(*[j[2] for j in jobs]):
        for flat, spec, img in jobs:
            call("nvo_mlp_tc_pack_weights", spec.desc, flat, img)
        ev = torch.cuda.Event()
        ev.record()
    for flat, spec, img in jobs:
        _prepacked[(flat.data_ptr(), id(spec))] = (img, ev)


def clear_prepacked() -> None:
    _prepacked.clear()
    _field_prepacked.clear()
    _prop_resident.clear()


def mlp_tc_forward(x16, wimage, spec: MlpSpec, n: int, save: bool, row_mask=None):
    """x16: TMH fp16 buffer of n rows; wimage from tc_pack_weights."""
    check(x16, "mlp_tc input", torch.float16, (tmh_numel(n, tc_in_pad(spec)),))
    check(wimage, "mlp_tc weight image", torch.uint8)
    if row_mask is not None:
        check(row_mask, "row_mask", torch.float32, (n,))
    y = torch.empty((n, spec.out_dim), dtype=torch.float32, device=x16.device)
    saved = None
    if save and len(spec.dims) > 1:
        saved = torch.empty(_tc_saved_bytes(spec, n), dtype=torch.uint8, device=x16.device)
    call("nvo_mlp_tc_forward", spec.desc, n, x16, wimage, row_mask, y, saved)
    return y, saved


# max|dy| handed from the kernel that produced a gradient to the tensor-core backward that consumes it: (data_ptr, numel) -> device float
# (bit pattern).  One entry at most, written by _FieldHeadsTC.backward and popped by the next mlp_tc_backward on that very tensor.
_dy_absmax: dict = {}


def mlp_tc_backward(x16, wimage, saved, y, dy, spec: MlpSpec, need_dx: bool, need_dparams: bool, dflat=None, row_mask=None, dy_absmax: float = 0.0):
    n = dy.shape[0]
    check(dy, "mlp dy", torch.float32, (n, spec.out_dim))
    # dx comes back in TMF layout ([tile][in_dim][128] fp32, see include/nvo_b200.h); tmf_to_rows() converts when needed
    dx = torch.empty(tmh_numel(n, spec.in_dim), dtype=torch.float32, device=x16.device) if need_dx else None
    if need_dparams and dflat is None:
        dflat = torch.zeros(spec.n_params, dtype=torch.float32, device=x16.device)
    hit = _dy_absmax.pop((dy.data_ptr(), dy.numel()), None) if dy_absmax == 0.0 else None
    scratch = None
    if hit is not None and hit[1].shape == dy.shape:
        scratch, dy_absmax = hit[0], -1.0  # `scratch` already holds max|dy|
    else:
        scratch = torch.empty(1, dtype=torch.float32, device=x16.device)
    call("nvo_mlp_tc_backward", spec.desc, n, x16, wimage, saved, y, row_mask, dy, float(dy_absmax), scratch, dx, dflat if need_dparams else None)
    return dx, dflat


def tmf_to_rows(src, n: int, k: int):
    check(src, "tmf buffer", torch.float32, (tmh_numel(n, k),))
    dst = torch.empty((n, k), dtype=torch.float32, device=src.device)
    call("nvo_tmf_to_rows", n, k, src, dst)
    return dst


def _flat_of(params):
    with torch.no_grad():
        flat = flat_alias(params)
        if flat is None:
            flat = torch.cat([p.reshape(-1).float() for p in params])
    return flat


def _split_grads(dflat, spec: MlpSpec):
    grads = []
    for (a, b), (ws, bs) in zip(_pairs(spec.offsets()), spec.shapes):
        grads.append(dflat[a[0]:a[1]].view(ws))
        grads.append(dflat[b[0]:b[1]].view(bs))
    return grads


class _MlpApplyTC(torch.autograd.Function):
    """fp32 [n,in_dim] input, cast + padded internally (tcnn.Network drop-in on the tensor-core path)."""

    @staticmethod
    def forward(ctx, x, spec, row_mask, *params):
        n = x.shape[0]
        x16 = cast_pad_f16(x.contiguous(), spec)
        wimage = tc_pack_weights(_flat_of(params), spec)
        y, saved = mlp_tc_forward(x16, wimage, spec, n, any(ctx.needs_input_grad), row_mask)
        ctx.save_for_backward(x16, wimage, saved, y, row_mask)
        ctx.spec, ctx.n_tensors = spec, len(params)
        ctx.main_grad = getattr(params[0], "_nvo_main_grad", None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x16, wimage, saved, y, row_mask = ctx.saved_tensors
        need_dx, need_dp = ctx.needs_input_grad[0], any(ctx.needs_input_grad[3:])
        dx, dflat = mlp_tc_backward(x16, wimage, saved, y, dy.contiguous(), ctx.spec, need_dx, need_dp, ctx.main_grad, row_mask)
        if dx is not None:
            dx = tmf_to_rows(dx, dy.shape[0], ctx.spec.in_dim)
        grads = [None] * ctx.n_tensors
        if need_dp and ctx.main_grad is None:
            grads = _split_grads(dflat, ctx.spec)
        return (dx, None, None, *grads)


def mlp_apply_tc(x, spec: MlpSpec, params, row_mask=None):
    return _MlpApplyTC.apply(x, spec, row_mask, *params)


class _GridMlpTC(torch.autograd.Function):
    """MLPWithHashEncoding on the tensor-core path: hash-grid features are produced directly in the MLP's fp16 operand
    layout and never cross autograd; x [n,3] fp32 -> y [n,out] fp32.  `cache` (a dict) receives the internal buffers
    (feat16, wimage, saved, y) so get_normals can run the input-gradient pass without recomputing the forward."""

    @staticmethod
    def forward(ctx, x, table, gspec, mspec, cache, *params):
        x = x.contiguous()
        n = x.shape[0]
        jac = None
        if cache is not None and cache.get("want_jac"):
            feat16, jac = grid_forward_jac(x, table, gspec)  # normals follow: keep d(feature)/dx instead of re-gathering the table
        else:
            feat16 = grid_forward(x, table, gspec, "tmh")
        wimage = tc_pack_weights(_flat_of(params), mspec)
        y, saved = mlp_tc_forward(feat16, wimage, mspec, n, any(ctx.needs_input_grad) or cache is not None)
        if cache is not None:
            cache["jac"] = jac
            # y.detach(): a separate tensor object — the returned `y` gets a grad_fn attached, and caching it would pin this step's
            # autograd graph (and its stream) into the next step, which breaks CUDA-graph capture
            cache.update(feat16=feat16, wimage=wimage, saved=saved, y=y.detach())
        ctx.save_for_backward(x, table, feat16, wimage, saved, y)
        ctx.gspec, ctx.mspec, ctx.n_tensors = gspec, mspec, len(params)
        ctx.table_main_grad = getattr(table, "_nvo_main_grad", None)
        ctx.mlp_main_grad = getattr(params[0], "_nvo_main_grad", None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, table, feat16, wimage, saved, y = ctx.saved_tensors
        need_dx, need_dt, need_dp = ctx.needs_input_grad[0], ctx.needs_input_grad[1], any(ctx.needs_input_grad[5:])
        dfeat, dflat = mlp_tc_backward(feat16, wimage, saved, y, dy.contiguous(), ctx.mspec, need_dx or need_dt, need_dp, ctx.mlp_main_grad)
        dtable = dx = None
        if need_dt:
            if ctx.table_main_grad is not None and leaf_streams.enabled:
                with leaf_streams.fork(x, dfeat):
                    grid_backward(x, dfeat, ctx.gspec, dtable=ctx.table_main_grad, tmf=True)
                    if leaf_streams.after_field_backward is not None and not need_dx:
                        leaf_streams.after_field_backward()
            elif ctx.table_main_grad is not None:
                grid_backward(x, dfeat, ctx.gspec, dtable=ctx.table_main_grad, tmf=True)
            else:
                dtable = grid_backward(x, dfeat, ctx.gspec, tmf=True).view(table.shape)
        if need_dx:
            dx = grid_backward_input(x, table, dfeat, ctx.gspec, tmf=True)
        grads = [None] * ctx.n_tensors
        if need_dp and ctx.mlp_main_grad is None:
            grads = _split_grads(dflat, ctx.mspec)
        return (dx, dtable, None, None, None, *grads)


def grid_mlp_tc(x, table, gspec: GridSpec, mspec: MlpSpec, params, cache=None):
    return _GridMlpTC.apply(x, table, gspec, mspec, cache, *params)


class _FieldHeadsTC(torch.autograd.Function):
    """NerfactoField.get_outputs on the tensor-core path: input assembly (fp16, padded) -> mlp_head -> rgb and
    mlp_pred_normals(+head, tanh) -> normalize -> pred_normals, plus density = trunc_exp(h0)*selector.
    Returns density [n], rgb [n,3], pred_normals [n,3] | None."""

    @staticmethod
    def forward(ctx, h, embedding, selector, directions, positions, cam_idx, B, S, head_spec, pn_spec, n_head, *params):
        ctx.set_materialize_grads(False)
        n = B * S
        dev = h.device
        h = check(h.contiguous(), "mlp_base output", torch.float32, (n, 16))
        embedding = check(embedding.contiguous(), "appearance embedding", torch.float32)
        head_params, pn_params = params[:n_head], params[n_head:]
        want_pn = pn_spec is not None
        density = torch.empty(n, dtype=torch.float32, device=dev)
        head_in = torch.empty(tmh_numel(n, 64), dtype=torch.float16, device=dev)
        pn_in = torch.empty(tmh_numel(n, 32), dtype=torch.float16, device=dev) if want_pn else None
        call("nvo_field_assemble_forward", B, S, h, selector, directions, positions, cam_idx, embedding, 1, density, head_in, pn_in)
        need = any(ctx.needs_input_grad)
        pn_flat = pn_saved = pn_raw = pn = None

        def pn_branch():
            flat = tc_pack_weights(_flat_of(pn_params), pn_spec)
            raw, sv = mlp_tc_forward(pn_in, flat, pn_spec, n, need)
            out = torch.empty_like(raw)
            call("nvo_normalize3_forward", n, raw, 1.0, 1e-12, out)
            return flat, raw, sv, out

        head_flat = tc_pack_weights(_flat_of(head_params), head_spec)
        if want_pn and leaf_streams.enabled and env_flag("NVO_FIELD_BRANCHES", True):
            # the predicted-normals network only shares its input assembly with the colour head: issue it on its own stream
            main = torch.cuda.current_stream()
            side = leaf_streams.branch_streams[1]
            side.wait_stream(main)
            with torch.cuda.stream(side):
                pn_flat, pn_raw, pn_saved, pn = pn_branch()
            rgb, head_saved = mlp_tc_forward(head_in, head_flat, head_spec, n, need)
            main.wait_stream(side)
            for t in (pn_flat, pn_raw, pn_saved, pn):
                if t is not None:
                    t.record_stream(main)  # allocated on the side stream, consumed (render, backward) on the main one
        else:
            rgb, head_saved = mlp_tc_forward(head_in, head_flat, head_spec, n, need)
            if want_pn:
                pn_flat, pn_raw, pn_saved, pn = pn_branch()
        ctx.save_for_backward(h, selector, cam_idx, head_in, head_flat, head_saved, rgb, pn_in, pn_flat, pn_saved, pn_raw)
        ctx.B, ctx.S, ctx.emb_shape, ctx.head_spec, ctx.pn_spec = B, S, embedding.shape, head_spec, pn_spec
        ctx.n_head, ctx.n_pn = len(head_params), len(pn_params)
        ctx.emb_main_grad = getattr(embedding, "_nvo_main_grad", None)
        ctx.head_main_grad = getattr(head_params[0], "_nvo_main_grad", None)
        ctx.pn_main_grad = getattr(pn_params[0], "_nvo_main_grad", None) if want_pn else None
        return density, rgb, pn

    @staticmethod
    def backward(ctx, ddensity, drgb, dpn):
        h, selector, cam_idx, head_in, head_flat, head_saved, rgb, pn_in, pn_flat, pn_saved, pn_raw = ctx.saved_tensors
        n, dev = h.shape[0], h.device
        c = lambda t: None if t is None else t.contiguous()
        if drgb is None:
            drgb = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        dhead_in, dhead_flat = mlp_tc_backward(head_in, head_flat, head_saved, rgb, drgb.contiguous(), ctx.head_spec, True, True, ctx.head_main_grad)
        dpn_in = dpn_flat = None
        if ctx.pn_spec is not None and dpn is not None:
            dpn_raw = torch.empty_like(pn_raw)
            call("nvo_normalize3_backward", n, pn_raw, dpn.contiguous(), 1.0, 1e-12, dpn_raw)
            dpn_in, dpn_flat = mlp_tc_backward(pn_in, pn_flat, pn_saved, pn_raw, dpn_raw, ctx.pn_spec, True, True, ctx.pn_main_grad)
        dh = torch.empty_like(h)
        demb = None
        if ctx.needs_input_grad[1]:
            if ctx.emb_main_grad is not None and cam_idx is not None:
                demb = ctx.emb_main_grad
            elif cam_idx is not None:
                demb = torch.zeros(ctx.emb_shape, dtype=torch.float32, device=dev)
            else:
                demb = tmf_to_rows(dhead_in, n, 63)[:, 31:].sum(0).reshape(ctx.emb_shape)
        # max|dh| is reduced inside this kernel and handed to the base network's backward (its gradient-scale pass disappears)
        amax = torch.zeros(1, dtype=torch.float32, device=dev)
        call("nvo_field_assemble_backward", ctx.B, ctx.S, h, selector, cam_idx, c(ddensity), dhead_in, dpn_in, 1, dh, demb if cam_idx is not None else None, amax)
        _dy_absmax.clear()
        _dy_absmax[(dh.data_ptr(), dh.numel())] = (amax, dh)  # dh is kept alive: its address cannot be handed to another tensor meanwhile
        if demb is ctx.emb_main_grad:
            demb = None
        head_grads = [None] * ctx.n_head
        if ctx.head_main_grad is None:
            head_grads = _split_grads(dhead_flat, ctx.head_spec)
        pn_grads = [None] * ctx.n_pn
        if ctx.pn_spec is not None and ctx.pn_main_grad is None:
            pn_grads = _split_grads(dpn_flat, ctx.pn_spec) if dpn_flat is not None else [None] * ctx.n_pn
        return (dh, demb, None, None, None, None, None, None, None, None, None, *head_grads, *pn_grads)


def field_heads_tc(h, embedding, selector, directions, positions, cam_idx, B, S, head_spec, head_params, pn_spec=None, pn_params=()):
    return _FieldHeadsTC.apply(h, embedding, selector, directions, positions, cam_idx, B, S, head_spec, pn_spec, len(head_params), *head_params, *pn_params)


# ------------------------------------------------------------------------------------------------
# fused field networks (csrc/field_tc.cu): mlp_base + normals chain + input assembly + mlp_head + mlp_pred_normals in one kernel
# ------------------------------------------------------------------------------------------------


_field_prepacked: dict = {}  # mlp_base parameter pointer -> (weight image, ready event), see prepack_field_weights


def field_pack_weights(gspec: GridSpec, base_flat, head_flat, pn_flat=None):
    """fp32 torch-layout parameters of the three field networks -> the fused kernels' fp16 weight image (uint8 tensor)."""
    check(base_flat, "mlp_base params", torch.float32, (32 * 64 + 64 + 64 * 16 + 16,))
    check(head_flat, "mlp_head params", torch.float32, (63 * 64 + 64 + 64 * 64 + 64 + 64 * 3 + 3,))
    if pn_flat is not None:
        check(pn_flat, "mlp_pred_normals (+ head) params", torch.float32, (27 * 64 + 64 + 2 * (64 * 64 + 64) + 64 * 3 + 3,))
    hit = _field_prepacked.get(base_flat.data_ptr())
    if hit is not None:
        torch.cuda.current_stream().wait_event(hit[1])
        return hit[0]
    img = torch.empty(int(_lib.load().nvo_field_wimage_bytes()), dtype=torch.uint8, device=base_flat.device)
    call("nvo_field_pack_weights", gspec.desc(torch.float32, "tmh"), base_flat, head_flat, pn_flat, img)
    return img


def prepack_field_weights(gspec: GridSpec, base_params, head_params, pn_params=None) -> None:
    """The fused field kernels' weight image packed ahead of its use on a side stream (the parameters only change in the optimizer);
    field_pack_weights() hands it out after waiting for the ready event.  Cleared by clear_prepacked()."""
    base_flat, head_flat = _flat_of(base_params), _flat_of(head_params)
    pn_flat = _flat_of(pn_params) if pn_params else None
    img = torch.empty(int(_lib.load().nvo_field_wimage_bytes()), dtype=torch.uint8, device=base_flat.device)
    with leaf_streams.fork(img):
        call("nvo_field_pack_weights", gspec.desc(torch.float32, "tmh"), base_flat, head_flat, pn_flat, img)
        ev = torch.cuda.Event()
        ev.record()
    _field_prepacked[base_flat.data_ptr()] = (img, ev)


def field_fused_supported(gspec: GridSpec, base_spec: MlpSpec, head_spec: MlpSpec, pn_spec: Optional[MlpSpec], S: int) -> bool:
    """The fused kernels are specialised for nerfacto's default field: 16 x 2 hash features -> 64 -> 16, head 63 -> 64 -> 64 -> 3,
    pred-normals 27 -> 64 -> 64 -> 64 -> 3 (csrc/field_tc.cu); at least 32 samples per ray (the backward's per-ray embedding reduction)."""
    ok = gspec.n_levels == 16 and gspec.features_per_level == 2 and base_spec.in_dim == 32 and tuple(base_spec.dims) == (64, 16)
    ok = ok and head_spec.in_dim == 63 and tuple(head_spec.dims) == (64, 64, 3)
    if pn_spec is not None:
        ok = ok and pn_spec.in_dim == 27 and tuple(pn_spec.dims) == (64, 64, 64, 3)
    return ok and S >= 32 and not env_flag("NVO_FIELD_PER_NETWORK", False)


def field_forward(feat16, jac, positions, directions, cam_idx, embedding, selector, wimage, B: int, S: int, want_pn: bool, save: bool, save_pn: bool = False):
    """One launch: (density [n], rgb [n,3], pred_normals [n,3] | None, normals [n,3] | None, h0 [n], pn_raw [n,3] | None, saved | None).
    jac (the grid forward's saved derivatives) enables the density-gradient normals."""
    n = B * S
    dev = feat16.device
    check(feat16, "hash features (tmh)", torch.float16, (tmh_numel(n, 32),))
    check(directions, "directions", torch.float32, (B, 3))
    check(selector, "selector", torch.float32, (n,))
    check(embedding, "appearance embedding", torch.float32)
    if cam_idx is not None:
        check(cam_idx, "camera indices", torch.int64, (B,))
    elif embedding.numel() != 32:
        raise RuntimeError("field_forward: without camera indices the embedding must be one 32-vector")
    if want_pn:
        check(positions, "positions", torch.float32, (n, 3))
    if jac is not None:
        check(jac, "grid jacobian", torch.float16, (3 * tmh_numel(n, 32),))
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    density, rgb, h0 = f(n), f(n, 3), f(n)
    pn = f(n, 3) if want_pn else None
    pn_raw = f(n, 3) if want_pn else None
    normals = f(n, 3) if jac is not None else None
    saved = torch.empty(int(_lib.load().nvo_field_saved_bytes(n, int(save_pn))), dtype=torch.uint8, device=dev) if save else None
    call("nvo_field_forward", B, S, feat16, jac, positions, directions, cam_idx, embedding, selector, wimage, density, rgb, pn, normals, h0, pn_raw, saved,
         int(save_pn))
    return density, rgb, pn, normals, h0, pn_raw, saved


FS_CHUNKS, FS_CHUNKS_HEAD, FS_P, FS_AP1, CHUNK_B = 60, 32, 32, 36, 2048  # saved-tile layout of csrc/field_tc.cu


def field_backward(feat16, saved, save_pn: bool, wimage, rgb, h0, selector, cam_idx, ddensity, drgb, dpn_in, B: int, S: int, dbase, dhead, demb,
                   directions=None, ddirections=None):
    """One launch: returns dfeat (fp32 TMF [tiles][32][128]); ACCUMULATES into dbase / dhead / demb (and ddirections [B,3])."""
    n = B * S
    check(drgb, "drgb", torch.float32, (n, 3))
    if ddensity is not None:
        check(ddensity, "ddensity", torch.float32, (n,))
    dfeat = torch.empty(tmh_numel(n, 32), dtype=torch.float32, device=feat16.device)
    scratch = torch.empty(2, dtype=torch.float32, device=feat16.device)
    call("nvo_field_backward", B, S, feat16, saved, int(save_pn), wimage, rgb, h0, selector, cam_idx, ddensity, drgb, dpn_in, scratch, dfeat, dbase, dhead, demb,
         directions, ddirections)
    return dfeat


# d loss / d (ray origins, ray directions) of the step in flight: (d_origins [B,3], d_directions [B,3]) zeroed by the trainer, or None.  While
# set, the field / proposal-density backward passes ACCUMULATE their contributions there (on whatever side stream they run on) instead of
# returning them through autograd; the trainer turns the sums into d loss / d pose_adjustment after joining the streams
# (CameraOptimizer.apply_to_raybundle's backward, NS/cameras/camera_optimizers.py:142-147).
ray_grad_sink = None


def position_backward(origins, directions, iv: "Intervals", dx, d_origins, d_directions) -> None:
    """Accumulates d loss / d (origins, directions) [B,3] from dx [B*S,3] = d loss / d (normalised contracted sample positions)."""
    s, e, stride = iv.triple()
    check(dx, "dx", torch.float32, (iv.B * iv.S, 3))
    call("nvo_position_backward", iv.B, iv.S, origins, directions, s, e, stride, dx, d_origins, d_directions)


class _FieldFused(torch.autograd.Function):
    """NerfactoField.forward on the fused tensor-core path (csrc/field_tc.cu): sample positions -> contraction -> hash grid (+ saved d feature / dx)
    -> ONE kernel for mlp_base, density-gradient normals, input assembly, mlp_head and mlp_pred_normals; backward: ONE kernel for mlp_head,
    assembly and mlp_base (+ the per-network kernel on the saved tiles for mlp_pred_normals when it receives a gradient), then the table
    scatter and, for camera-pose optimisation, d loss / d (origins, directions).
    Returns density [n], rgb [n,3], pred_normals [n,3] | None, normals [n,3] | None (no gradient: base_field.py:92-97 is first order),
    h0 [n] (raw density, no gradient), x [n,3] (normalised sample locations, no gradient)."""

    @staticmethod
    def forward(ctx, origins, directions, table, embedding, cam_idx, iv, B, S, gspec, want_normals, save_pn, pn_spec, n_base, n_head, *params):
        ctx.set_materialize_grads(False)
        n = B * S
        base_params, head_params, pn_params = params[:n_base], params[n_base:n_base + n_head], params[n_base + n_head:]
        want_pn = len(pn_params) > 0
        origins = check(origins.contiguous(), "origins", torch.float32, (B, 3))
        directions = check(directions.contiguous(), "directions", torch.float32, (B, 3))
        positions = torch.empty((n, 3), dtype=torch.float32, device=origins.device)
        x, selector = torch.empty_like(positions), torch.empty(n, dtype=torch.float32, device=origins.device)
        s_, e_, stride_ = iv.triple()
        call("nvo_sample_positions_contract", B, S, origins, directions, s_, e_, stride_, positions, x, selector)
        need = any(ctx.needs_input_grad)
        want_ray_grads = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or (ray_grad_sink is not None and need)
        if want_normals or want_ray_grads:
            feat16, jac = grid_forward_jac(x, table, gspec)
        else:
            feat16, jac = grid_forward(x, table, gspec, "tmh"), None
        wimage = field_pack_weights(gspec, _flat_of(base_params), _flat_of(head_params), _flat_of(pn_params) if want_pn else None)
        embedding = check(embedding.contiguous(), "appearance embedding", torch.float32)
        density, rgb, pn, normals, h0, pn_raw, saved = field_forward(feat16, jac if want_normals else None, positions if want_pn else None, directions, cam_idx,
                                                                     embedding, selector, wimage, B, S, want_pn, need, bool(save_pn and want_pn))
        # a trainable mlp_pred_normals runs its backward through the per-network kernel (its own weight-image format)
        pn_wimage = tc_pack_weights(_flat_of(pn_params), pn_spec) if (need and save_pn and want_pn) else None
        ctx.save_for_backward(x, table, feat16, wimage, saved, rgb, h0, selector, cam_idx, pn_raw, pn_wimage, origins, directions,
                              jac if want_ray_grads else None)
        ctx.B, ctx.S, ctx.gspec, ctx.pn_spec, ctx.save_pn, ctx.emb_shape = B, S, gspec, pn_spec, bool(save_pn and want_pn), embedding.shape
        ctx.iv, ctx.sink = iv, ray_grad_sink if want_ray_grads else None
        ctx.n_base, ctx.n_head, ctx.n_pn = n_base, n_head, len(pn_params)
        ctx.base_spec_shapes = [tuple(p.shape) for p in base_params]
        ctx.head_spec_shapes = [tuple(p.shape) for p in head_params]
        ctx.pn_shapes = [tuple(p.shape) for p in pn_params]
        ctx.table_main_grad = getattr(table, "_nvo_main_grad", None)
        ctx.emb_main_grad = getattr(embedding, "_nvo_main_grad", None)
        ctx.base_main_grad = getattr(base_params[0], "_nvo_main_grad", None)
        ctx.head_main_grad = getattr(head_params[0], "_nvo_main_grad", None)
        ctx.pn_main_grad = getattr(pn_params[0], "_nvo_main_grad", None) if want_pn else None
        nd = [t for t in (normals, h0, x) if t is not None]
        ctx.mark_non_differentiable(*nd)
        return density, rgb, pn, normals, h0, x

    @staticmethod
    def backward(ctx, ddensity, drgb, dpn, _dnormals, _dh0, _dx):
        x, table, feat16, wimage, saved, rgb, h0, selector, cam_idx, pn_raw, pn_wimage, origins, directions, jac = ctx.saved_tensors
        B, S = ctx.B, ctx.S
        n, dev = B * S, x.device
        c = lambda t: None if t is None else t.contiguous()
        if drgb is None:
            drgb = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        flat_numel = lambda shapes: sum(int(np.prod(s)) for s in shapes)
        zeros = lambda k: torch.zeros(k, dtype=torch.float32, device=dev)
        # mlp_pred_normals: Normalize' -> the per-network tensor-core backward reading its input / activations out of the fused saved tiles
        dpn_in = dpn_flat = None
        if dpn is not None and ctx.n_pn:
            if not ctx.save_pn:
                raise RuntimeError("fused field: pred_normals received a gradient but its activations were not saved (construct the field with "
                                   "pred_normals_trainable=True)")
            dpn_raw = torch.empty_like(pn_raw)
            call("nvo_normalize3_backward", n, pn_raw, dpn.contiguous(), 1.0, 1e-12, dpn_raw)
            spec = ctx.pn_spec
            dpn_in = torch.empty(tmh_numel(n, spec.in_dim), dtype=torch.float32, device=dev)
            dpn_flat = ctx.pn_main_grad if ctx.pn_main_grad is not None else zeros(spec.n_params)
            scratch = torch.empty(1, dtype=torch.float32, device=dev)
            tile_bytes = FS_CHUNKS * CHUNK_B
            call("nvo_mlp_tc_backward_strided", spec.desc, n, saved[FS_P * CHUNK_B:], tile_bytes, pn_wimage, saved[FS_AP1 * CHUNK_B:], tile_bytes, pn_raw, None,
                 dpn_raw, 0.0, scratch, dpn_in, dpn_flat)
        dbase = ctx.base_main_grad if ctx.base_main_grad is not None else zeros(flat_numel(ctx.base_spec_shapes))
        dhead = ctx.head_main_grad if ctx.head_main_grad is not None else zeros(flat_numel(ctx.head_spec_shapes))
        demb = None
        if ctx.needs_input_grad[3]:
            demb = ctx.emb_main_grad if (ctx.emb_main_grad is not None and cam_idx is not None) else torch.zeros(ctx.emb_shape, dtype=torch.float32, device=dev)
        # camera-pose optimisation: d loss / d (origins, directions), into the trainer's sink or back through autograd
        need_rays = jac is not None
        d_o = d_d = None
        if need_rays:
            d_o, d_d = ctx.sink if ctx.sink is not None else (torch.zeros((B, 3), dtype=torch.float32, device=dev), torch.zeros((B, 3), dtype=torch.float32, device=dev))
        dfeat = field_backward(feat16, saved, ctx.save_pn, wimage, rgb, h0, selector, cam_idx, c(ddensity), drgb.contiguous(), dpn_in, B, S, dbase, dhead, demb,
                               directions if need_rays else None, d_d)
        need_dt = ctx.needs_input_grad[2]
        dtable = None
        if need_dt:
            if ctx.table_main_grad is not None and leaf_streams.enabled:
                with leaf_streams.fork(x, dfeat, critical=True):
                    grid_backward(x, dfeat, ctx.gspec, dtable=ctx.table_main_grad, tmf=True)
                    if leaf_streams.after_field_backward is not None:
                        leaf_streams.after_field_backward()
            elif ctx.table_main_grad is not None:
                grid_backward(x, dfeat, ctx.gspec, dtable=ctx.table_main_grad, tmf=True)
            else:
                dtable = grid_backward(x, dfeat, ctx.gspec, tmf=True).view(table.shape)
        if need_rays:
            # d x through the saved feature derivatives (one streaming pass instead of a second gather over the table), then the contraction
            # Jacobian and the per-ray sums; positions feed mlp_pred_normals' encoding too, but that path carries no gradient here
            def rays():
                dxn = grid_jac_dx(jac, dfeat, ctx.gspec, n)
                position_backward(origins, directions, ctx.iv, dxn, d_o, d_d)

            if ctx.sink is not None and leaf_streams.enabled:
                with leaf_streams.fork(jac, dfeat, origins, directions, ctx.iv):
                    rays()
            else:
                rays()

        def split(flat, shapes, main):
            if main is not None or flat is None:
                return [None] * len(shapes)
            out, off = [], 0
            for s in shapes:
                k = int(np.prod(s))
                out.append(flat[off:off + k].view(s))
                off += k
            return out

        if demb is ctx.emb_main_grad:
            demb = None
        if ctx.sink is not None:
            d_o = d_d = None
        grads = split(dbase, ctx.base_spec_shapes, ctx.base_main_grad) + split(dhead, ctx.head_spec_shapes, ctx.head_main_grad) + split(dpn_flat, ctx.pn_shapes, ctx.pn_main_grad)
        return (d_o if ctx.needs_input_grad[0] else None, d_d if ctx.needs_input_grad[1] else None, dtable, demb, None, None, None, None, None, None, None, None,
                None, None, *grads)


def field_fused(origins, directions, iv, table, embedding, cam_idx, B: int, S: int, gspec: GridSpec, want_normals: bool, base_params, head_params,
                pn_spec: Optional[MlpSpec] = None, pn_params=(), save_pn: bool = True):
    """origins / directions [B,3] per ray, iv = the samples' Euclidean intervals; see _FieldFused."""
    return _FieldFused.apply(origins, directions, table, embedding, cam_idx, iv, B, S, gspec, want_normals, save_pn, pn_spec, len(base_params),
                             len(head_params), *base_params, *head_params, *pn_params)


# ------------------------------------------------------------------------------------------------
# step prologue (csrc/batch.cu): pixel sampling + gather + ray generation + camera-pose correction
# ------------------------------------------------------------------------------------------------


def _check_cameras(intrinsics, extrinsics):
    check(intrinsics, "camera intrinsics", torch.float32, (None, 4))
    check(extrinsics, "camera extrinsics", torch.float32, (None, 4, 4))


def batch_prologue(u, num_active: int, intrinsics, extrinsics, frames_color, frames_depth, frames_normal=None, pose_adjustment=None, pose_mode: int = 0,
                   want_raw_directions: bool = False, num_active_dev=None):
    """One launch of nvo_batch_prologue (include/nvo_b200.h): returns a dict with indices / camera_indices (int64), origins, directions,
    directions_norm, pixel_area, image, depth_image, normal_image (when frames_normal is given) [, directions_raw].
    num_active_dev: device int32 [1] holding the number of active frames; the kernel then reads the count from it (clamped to the frame
    capacity), so a CUDA graph captured around this call keeps following keyframe insertions."""
    check(u, "uniform draws", torch.float32, (None, 3))
    _check_cameras(intrinsics, extrinsics)
    K_alloc, H, W, _ = frames_color.shape
    check(frames_color, "frames_color", torch.float32, (K_alloc, H, W, 3))
    check(frames_depth, "frames_depth", torch.float32, (K_alloc, H, W, 1))
    if frames_normal is not None:
        check(frames_normal, "frames_normal", torch.float32, (K_alloc, H, W, 3))
    if not (0 < num_active <= K_alloc and num_active <= intrinsics.shape[0] and num_active <= extrinsics.shape[0]):
        raise RuntimeError(f"num_active={num_active} out of range for {K_alloc} frame slots / {intrinsics.shape[0]} cameras")
    if pose_mode != 0:
        check(pose_adjustment, "pose_adjustment", torch.float32, (None, 6))
        if pose_adjustment.shape[0] < num_active:
            raise RuntimeError("pose_adjustment has fewer rows than active cameras")
    B, dev = u.shape[0], u.device
    if num_active_dev is not None:
        check(num_active_dev, "num_active_dev", torch.int32, (1,))
        num_active = min(K_alloc, intrinsics.shape[0], extrinsics.shape[0], pose_adjustment.shape[0] if pose_mode != 0 else K_alloc)  # capacity bound
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    out = {"indices": torch.empty((B, 3), dtype=torch.int64, device=dev), "camera_indices": torch.empty((B, 1), dtype=torch.int64, device=dev),
           "origins": f(B, 3), "directions": f(B, 3), "directions_norm": f(B, 1), "pixel_area": f(B, 1), "image": f(B, 3), "depth_image": f(B, 1)}
    if frames_normal is not None:
        out["normal_image"] = f(B, 3)
    if want_raw_directions:
        out["directions_raw"] = f(B, 3)
    call("nvo_batch_prologue", B, num_active, num_active_dev, H, W, u, intrinsics, extrinsics, frames_color, frames_depth, frames_normal,
         pose_adjustment.detach() if pose_mode != 0 else None, pose_mode, out["indices"], out["camera_indices"], out["origins"], out["directions"],
         out["directions_norm"], out["pixel_area"], out["image"], out["depth_image"], out.get("normal_image"), out.get("directions_raw"))
    return out


def generate_rays(intrinsics, extrinsics, indices=None, cam: int = 0, height: int = 0, width: int = 0):
    """Pinhole rays for indices [n,3] (camera,row,col) or every pixel of frame `cam`: (origins, directions, directions_norm, pixel_area, camera_indices)."""
    _check_cameras(intrinsics, extrinsics)
    dev = intrinsics.device
    if indices is not None:
        check(indices, "ray indices", torch.int64, (None, 3))
        n = indices.shape[0]
    else:
        if not (0 <= cam < extrinsics.shape[0]) or height <= 0 or width <= 0:
            raise RuntimeError(f"generate_rays: camera {cam} / image size {height}x{width} invalid")
        n = height * width
    f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    o, d, dn, pa, ci = f(n, 3), f(n, 3), f(n, 1), f(n, 1), torch.empty((n, 1), dtype=torch.int64, device=dev)
    call("nvo_generate_rays", n, cam, width, indices, intrinsics, extrinsics, ci, o, d, dn, pa)
    return o, d, dn, pa, ci


class _PoseExpMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tangent, mode):
        out = torch.empty((tangent.shape[0], 3, 4), dtype=torch.float32, device=tangent.device)
        call("nvo_pose_exp_map", tangent.shape[0], mode, tangent, out)
        ctx.save_for_backward(tangent)
        ctx.mode = mode
        return out

    @staticmethod
    def backward(ctx, dM):
        (tangent,) = ctx.saved_tensors
        n = tangent.shape[0]
        d = torch.zeros_like(tangent)
        # the per-camera cotangent IS dM here: k_pose_grad alone (B = 0 rays), scratch = dM
        empty_i = torch.empty(0, dtype=torch.int64, device=tangent.device)
        empty_f = torch.empty(0, dtype=torch.float32, device=tangent.device)
        call("nvo_pose_correction_backward", 0, n, ctx.mode, empty_i, empty_f, empty_f, empty_f, tangent, dM.contiguous().view(n, 12), d)
        return d, None


def pose_exp_map(tangent, mode: int):
    check(tangent, "pose tangent", torch.float32, (None, 6))
    return _PoseExpMap.apply(tangent, mode)


class _PoseCorrection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, origins, directions, cam_idx, pose, mode):
        # no materialised zero gradients: under MappingTrainer's gradient sink nothing flows back through autograd, and the engine would
        # otherwise run this node's backward on zeros
        ctx.set_materialize_grads(False)
        o, d = torch.empty_like(origins), torch.empty_like(directions)
        call("nvo_pose_apply", origins.shape[0], mode, cam_idx, pose, origins, directions, o, d)
        ctx.save_for_backward(directions, cam_idx, pose)
        ctx.mode = mode
        return o, d

    @staticmethod
    def backward(ctx, do, dd):
        if do is None and dd is None:
            return None, None, None, None, None
        directions, cam_idx, pose = ctx.saved_tensors
        K, B = pose.shape[0], directions.shape[0]
        dpose = torch.zeros_like(pose)
        scratch = torch.zeros((K, 12), dtype=torch.float32, device=pose.device)
        do = torch.zeros_like(directions) if do is None else do.contiguous()
        dd = torch.zeros_like(directions) if dd is None else dd.contiguous()
        call("nvo_pose_correction_backward", B, K, ctx.mode, cam_idx, directions, do, dd, pose, scratch, dpose)
        d_dir = None
        if ctx.needs_input_grad[1]:  # R^T dd
            M = torch.empty((K, 3, 4), dtype=torch.float32, device=pose.device)
            call("nvo_pose_exp_map", K, ctx.mode, pose, M)
            d_dir = torch.bmm(M[cam_idx][:, :, :3].transpose(1, 2), dd[..., None]).squeeze(-1)
        return (do if ctx.needs_input_grad[0] else None), d_dir, None, dpose, None


def pose_correction(origins, directions, cam_idx, pose_adjustment, mode: int):
    """CameraOptimizer.apply_to_raybundle on an existing bundle (the fused prologue applies it in-kernel): origins + t, R @ directions,
    with the gradient w.r.t. pose_adjustment from nvo_pose_correction_backward."""
    check(origins, "origins", torch.float32, (None, 3))
    check(directions, "directions", torch.float32, (origins.shape[0], 3))
    check(cam_idx, "camera indices", torch.int64, (origins.shape[0],))
    check(pose_adjustment, "pose_adjustment", torch.float32, (None, 6))
    return _PoseCorrection.apply(origins, directions, cam_idx, pose_adjustment, mode)


def pose_correction_backward(cam_idx, directions_raw, d_origins, d_directions, pose_adjustment, mode: int, d_pose=None):
    """Gradient of the fused prologue's pose correction: d_pose[K,6] (+)= from dL/d(origins, directions) [B,3]."""
    K = pose_adjustment.shape[0]
    if d_pose is None:
        d_pose = torch.zeros_like(pose_adjustment)
    scratch = torch.zeros((K, 12), dtype=torch.float32, device=pose_adjustment.device)
    call("nvo_pose_correction_backward", directions_raw.shape[0], K, mode, cam_idx.reshape(-1), directions_raw, d_origins, d_directions, pose_adjustment, scratch, d_pose)
    return d_pose
