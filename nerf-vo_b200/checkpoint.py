"""Checkpoint / on-disk format compatibility.

The reference writes `step-%09d.ckpt` with torch.save (NS/engine/trainer.py:424-452):

    {"step": int,
     "pipeline":   pipeline.state_dict()            # keys "_model.<model key>" (+ "module." under DDP), datamanager keys
     "optimizers": {group: torch.optim.Adam.state_dict()},   # groups "fields", "proposal_networks", "camera_opt"
     "schedulers": {...}, "scalers": GradScaler.state_dict()}

and loads it back through Trainer._load_checkpoint / VanillaPipeline.load_pipeline / load_state_dict
(NS/engine/trainer.py:388-422, NS/pipelines/base_pipeline.py:110-135,412-423: strips "module." and "_model.").  NeRF-VO also
stores the keyframe store as `dataset.pt` (nerf_vo/mapping/nerfstudio_utils.py:230-241).

Two parameter layouts exist in reference checkpoints, depending on which implementation trained them:
  * "torch"  (implementation="torch"): `...hash_table`, `...layers.{i}.{weight,bias}` — the layout of this repo's modules: keys, shapes
    and dtypes are identical (tests/test_checkpoint.py checks them against a checkpoint written by the unmodified reference);
  * "tcnn"   (implementation="tcnn", NeRF-VO's default): one flat fp32 `params` tensor per tinycudann module —
    NetworkWithInputEncoding: [network | encoding] (TCNN/include/tiny-cuda-nn/network_with_input_encoding.h:115-134),
    FullyFusedMLP: row-major [out, in] matrices, input padded to 16 columns, output to 16 rows, no biases
    (TCNN/src/fully_fused_mlp.cu:656-669), GridEncoding: levels back to back, a level holds min(res^3 rounded up to 8, 2^log2_T) rows
    (TCNN/include/tiny-cuda-nn/encodings/grid.h:692-723).  `unpack_tcnn_*` below split such tensors; see `convert_tcnn_state` for what
    can and cannot be carried over (tiny-cuda-nn evaluates a different grid geometry than the torch path: SURVEY §8 a-notes).

Pure host-side tensor plumbing: no kernels, runs on CPU tensors as well as CUDA ones."""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch

GROUPS = ("fields", "proposal_networks", "camera_opt")


# ---------------------------------------------------------------------------------------------------------------------
# key handling
# ---------------------------------------------------------------------------------------------------------------------
def split_pipeline_state(pipeline_state: Dict[str, torch.Tensor]) -> Tuple["OrderedDict[str, torch.Tensor]", "OrderedDict[str, torch.Tensor]"]:
    """(model_state, other_state) with the "module." (DDP) and "_model." prefixes removed, order preserved
    (NS/pipelines/base_pipeline.py:110-135,418-421)."""
    model, other = OrderedDict(), OrderedDict()
    for k, v in pipeline_state.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if k.startswith("_model."):
            k = k[len("_model."):]
            if k.startswith("module."):
                k = k[len("module."):]
            model[k] = v
        else:
            other[k] = v
    return model, other


def layout_of(model_state: Dict[str, torch.Tensor]) -> str:
    """'tcnn' if the state holds tinycudann flat `params` tensors, else 'torch'."""
    return "tcnn" if any(k.endswith(".params") and (".tcnn_encoding." in k or ".model." in k) for k in model_state) else "torch"


def _group_of(key: str) -> Optional[str]:
    if key.startswith("proposal_networks."):
        return "proposal_networks"
    if key.startswith("field."):
        return "fields"
    if key.startswith("camera_optimizer."):
        return "camera_opt"
    return None


# ---------------------------------------------------------------------------------------------------------------------
# load
# ---------------------------------------------------------------------------------------------------------------------
def load_checkpoint(ckpt, model, camera_optimizer=None, trainer=None, strict: bool = True, map_location="cpu") -> Dict:
    """Loads a reference (or own) checkpoint — a path or the already loaded dict — into `model` (and, when given, the camera optimizer
    and the MappingTrainer's Adam state).  Returns {"step", "layout", "missing", "unexpected"}.

    strict=True raises on model keys of the field / proposal networks that are missing or have another shape (Module.load_state_dict
    semantics, NS/pipelines/base_pipeline.py:128); keys of modules this repo does not have (lpips weights, device_indicator_param,
    datamanager state) are reported in "unexpected" and ignored, as the reference's strict=False fallback does."""
    if not isinstance(ckpt, dict):
        ckpt = torch.load(ckpt, map_location=map_location, weights_only=False)
    if trainer is not None and hasattr(trainer, "_pending_fields"):
        trainer._pending_fields = False  # a deferred update of the state being replaced is dropped with it
    pipeline_state = ckpt["pipeline"] if "pipeline" in ckpt else ckpt
    model_state, other = split_pipeline_state(pipeline_state)
    layout = layout_of(model_state)
    if layout == "tcnn":
        raise NotImplementedError(
            "this checkpoint was trained with implementation='tcnn': its flat `params` tensors can be split with checkpoint.convert_tcnn_state(), "
            "but tiny-cuda-nn's grid geometry (scale 2^(l*log2 s)*N-1, +0.5 offset, dense coarse levels, bias-free MLPs) is not the torch path's "
            "that these kernels reproduce, so it cannot be loaded as-is")
    own = model.state_dict()
    missing = [k for k in own if k not in model_state]
    unexpected = [k for k in model_state if k not in own and not k.startswith("camera_optimizer.")]
    bad_shape = [k for k in own if k in model_state and tuple(own[k].shape) != tuple(model_state[k].shape)]
    if strict and (missing or bad_shape):
        raise RuntimeError(f"checkpoint does not match the model: missing {missing[:5]}{'...' if len(missing) > 5 else ''}, "
                           f"shape mismatch {[(k, tuple(model_state[k].shape), tuple(own[k].shape)) for k in bad_shape[:5]]}")
    with torch.no_grad():
        for k, dst in own.items():
            if k in model_state and k not in bad_shape:
                dst.copy_(model_state[k].to(dst.device, dst.dtype))  # in place: parameters may be views of a trainer's flat buffer
    if camera_optimizer is not None and "camera_optimizer.pose_adjustment" in model_state and hasattr(camera_optimizer, "pose_adjustment"):
        src = model_state["camera_optimizer.pose_adjustment"]
        with torch.no_grad():
            n = min(src.shape[0], camera_optimizer.pose_adjustment.shape[0])
            camera_optimizer.pose_adjustment[:n].copy_(src[:n].to(camera_optimizer.pose_adjustment.device))
    step = int(ckpt.get("step", 0)) if isinstance(ckpt, dict) else 0
    if trainer is not None and "optimizers" in ckpt:
        load_optimizer_state(ckpt["optimizers"], model_state, model, trainer)
        trainer.iteration = step + 1  # NS/engine/trainer.py:402: training resumes at the step after the saved one
    if hasattr(model, "proposal_sampler"):
        model.proposal_sampler._step = step
    return {"step": step, "layout": layout, "missing": missing, "unexpected": unexpected + list(other)}


def _ordered_param_keys(model_state_keys, model, group: str) -> List[str]:
    """Parameter keys of one optimizer group in the order of the group's parameter list — module registration order, which is the order of
    the state dict (NS/models/nerfacto.py:244-249: list(self.field.parameters()), list(self.proposal_networks.parameters()))."""
    params = {n for n, _ in model.named_parameters()}
    return [k for k in model_state_keys if k in params and _group_of(k) == group]


def load_optimizer_state(optimizers: Dict, model_state, model, trainer) -> None:
    """torch.optim.Adam.state_dict() per group -> the trainer's flat moment buffers and per-group step counters (every trainer arm: the fused
    peer arm keeps its own slice of the assembled whole-buffer moments, MappingTrainer.load_full_moments)."""
    views = {id(p): v for p, v in zip(trainer.params, trainer._views)}
    named = dict(model.named_parameters())
    fused = trainer.exp_avg is None
    if fused:
        exp_avg, exp_avg_sq = trainer.full_moments()  # collective in the fused arm: every rank loads the checkpoint
    else:
        exp_avg, exp_avg_sq = trainer.exp_avg, trainer.exp_avg_sq
    for gi, (gname, _, _) in enumerate(trainer.groups):
        if gname not in optimizers:
            continue
        sd = optimizers[gname]
        keys = _ordered_param_keys(list(model_state.keys()), model, gname)
        steps = []
        for idx, k in enumerate(keys):
            st = sd["state"].get(idx)
            if st is None or id(named[k]) not in views:  # a parameter that never received a gradient has no state (torch skips grad=None)
                continue
            off, n = views[id(named[k])]
            exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1).to(exp_avg.device))
            exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1).to(exp_avg.device))
            steps.append(int(st["step"]))
        if steps:
            # torch keeps `step` per parameter; every parameter of a group that receives gradients is stepped together
            trainer.step_counts[gi].fill_(max(steps))
    if fused:
        trainer.load_full_moments(exp_avg, exp_avg_sq)
        trainer.peer.reset_epochs()  # the step counters may have gone back: the exchange barriers' epoch flags follow (collective)


# ---------------------------------------------------------------------------------------------------------------------
# save
# ---------------------------------------------------------------------------------------------------------------------
def checkpoint_dict(step: int, model, camera_optimizer=None, trainer=None) -> Dict:
    """The dict Trainer.save_checkpoint writes (NS/engine/trainer.py:436-447), torch key layout."""
    if trainer is not None and hasattr(trainer, "flush"):
        trainer.flush()  # defer_fields_update: the last step's fields-group update is applied before the state is read
    pipeline = OrderedDict()
    # Model.device_indicator_param (NS/models/base_model.py:81): an empty parameter the reference's strict load_state_dict expects first
    pipeline["_model.device_indicator_param"] = torch.empty(0)
    for k, v in model.state_dict().items():
        pipeline["_model." + k] = v.detach().clone().cpu()
    if camera_optimizer is not None and hasattr(camera_optimizer, "pose_adjustment"):
        pipeline["_model.camera_optimizer.pose_adjustment"] = camera_optimizer.pose_adjustment.detach().clone().cpu()
    # "scalers": NeRF-VO trains with mixed_precision=True, so the reference's GradScaler is enabled and Trainer._load_checkpoint calls
    # grad_scaler.load_state_dict(loaded_state["scalers"]) (NS/engine/trainer.py:416), which raises on an empty dict.  These kernels scale
    # gradients per launch on the device (csrc/mlp_tc.cu) and keep no global loss scale: a fresh GradScaler's state is written.
    scalers = {"scale": 65536.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 0}
    out = {"step": int(step), "pipeline": pipeline, "optimizers": {}, "schedulers": {}, "scalers": scalers}
    if trainer is not None:
        exp_avg, exp_avg_sq = trainer.full_moments()  # fused peer arm: a collective (every rank calls checkpoint_dict / save_checkpoint)
        views = {id(p): v for p, v in zip(trainer.params, trainer._views)}
        named = dict(model.named_parameters())
        keys_all = list(model.state_dict().keys())
        for gi, (gname, _, _) in enumerate(trainer.groups):
            keys = _ordered_param_keys(keys_all, model, gname)
            state = {}
            t = int(trainer.step_counts[gi])
            for idx, k in enumerate(keys):
                off, n = views[id(named[k])]
                if t == 0:
                    continue
                state[idx] = {"step": torch.tensor(float(t)), "exp_avg": exp_avg[off:off + n].view(named[k].shape).clone().cpu(),
                              "exp_avg_sq": exp_avg_sq[off:off + n].view(named[k].shape).clone().cpu()}
            group = {"lr": trainer.lr, "betas": tuple(trainer.betas), "eps": trainer.eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
                     "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": list(range(len(keys)))}
            out["optimizers"][gname] = {"state": state, "param_groups": [group]}
        if getattr(trainer, "cam_group", None) is not None:
            # the "camera_opt" group (Adam 1e-4, ExponentialDecayScheduler; nerf_vo/mapping/nerfstudio.py:93-100) and its LambdaLR state
            import math

            _, off, _ = trainer.cam_group
            pose = model.camera_optimizer.pose_adjustment
            t = int(trainer.cam_step)
            frac = min(max(t / trainer.max_num_iterations, 0.0), 1.0)
            lr_t = math.exp(math.log(trainer.cam_lr) * (1 - frac) + math.log(trainer.cam_lr_final) * frac)
            state = {}
            if t > 0:
                state[0] = {"step": torch.tensor(float(t)), "exp_avg": exp_avg[off:off + pose.numel()].view(pose.shape).clone().cpu(),
                            "exp_avg_sq": exp_avg_sq[off:off + pose.numel()].view(pose.shape).clone().cpu()}
            group = {"lr": lr_t, "initial_lr": trainer.cam_lr, "betas": tuple(trainer.betas), "eps": trainer.eps, "weight_decay": 0, "amsgrad": False,
                     "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": [0]}
            out["optimizers"]["camera_opt"] = {"state": state, "param_groups": [group]}
            out["schedulers"]["camera_opt"] = {"base_lrs": [trainer.cam_lr], "last_epoch": t, "verbose": False, "_step_count": t + 1,
                                               "_get_lr_called_within_step": False, "_last_lr": [lr_t], "lr_lambdas": [None]}
    return out


def save_checkpoint(path: str, step: int, model, camera_optimizer=None, trainer=None) -> None:
    torch.save(checkpoint_dict(step, model, camera_optimizer, trainer), path)


def checkpoint_name(step: int) -> str:
    return f"step-{step:09d}.ckpt"


# ---------------------------------------------------------------------------------------------------------------------
# dataset.pt (nerf_vo/mapping/nerfstudio_utils.py:230-241)
# ---------------------------------------------------------------------------------------------------------------------
def save_dataset(dataset, path: str) -> None:
    k = dataset.num_active_frames
    out = {"camera_intrinsics": dataset.camera_intrinsics, "camera_extrinsics": dataset.camera_extrinsics[:k], "frames_color": dataset.frames_color[:k],
           "frames_depth": dataset.frames_depth[:k]}
    if getattr(dataset, "use_normals", False):
        out["frames_normal"] = dataset.frames_normal[:k]
    torch.save(out, path)


def load_dataset(dataset, path_or_dict) -> int:
    """Fills a DynamicDataset from a reference `dataset.pt`; returns the number of active frames."""
    d = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu", weights_only=False)
    k = d["frames_color"].shape[0]
    if k > dataset.frames_color.shape[0]:
        raise RuntimeError(f"dataset.pt holds {k} frames, the store was built for {dataset.frames_color.shape[0]}")
    with torch.no_grad():
        ci = d["camera_intrinsics"]
        if torch.is_tensor(ci) and torch.is_tensor(dataset.camera_intrinsics) and ci.shape == dataset.camera_intrinsics.shape:
            dataset.camera_intrinsics.copy_(ci.to(dataset.camera_intrinsics.device))
        dataset.camera_extrinsics[:k].copy_(d["camera_extrinsics"].to(dataset.camera_extrinsics.device))
        dataset.frames_color[:k].copy_(d["frames_color"].to(dataset.frames_color.device))
        dataset.frames_depth[:k].copy_(d["frames_depth"].to(dataset.frames_depth.device))
        if "frames_normal" in d and getattr(dataset, "use_normals", False):
            dataset.frames_normal[:k].copy_(d["frames_normal"].to(dataset.frames_normal.device))
    dataset.num_active_frames = k
    return k


# ---------------------------------------------------------------------------------------------------------------------
# tiny-cuda-nn flat-parameter layouts
# ---------------------------------------------------------------------------------------------------------------------
def tcnn_grid_levels(n_levels: int, base_resolution: int, per_level_scale: float, log2_hashmap_size: int) -> List[Dict]:
    """Per level: {"offset": first row, "rows", "resolution", "scale", "hashed"} exactly as GridEncoding's constructor lays the levels out
    (grid.h:692-723; scale / resolution: common_device.h:709-718, computed in fp32 like the C++)."""
    import numpy as np

    out, offset = [], 0
    log2_s = np.float32(math.log2(per_level_scale))
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_s)) * np.float32(base_resolution) - np.float32(1.0)
        res = int(math.ceil(float(scale))) + 1
        dense = res ** 3
        rows = min((min(dense, (2 ** 32 - 1) // 2) + 7) // 8 * 8, 1 << log2_hashmap_size)
        out.append({"offset": offset, "rows": rows, "resolution": res, "scale": float(scale), "hashed": dense > rows})
        offset += rows
    return out


def tcnn_mlp_n_params(in_dim: int, width: int, n_hidden_layers: int, out_dim: int) -> int:
    in_pad, out_pad = (in_dim + 15) // 16 * 16, (out_dim + 15) // 16 * 16
    return width * in_pad + (n_hidden_layers - 1) * width * width + out_pad * width


def unpack_tcnn_mlp(params: torch.Tensor, in_dim: int, width: int, n_hidden_layers: int, out_dim: int, ones_padded_input: bool = True):
    """FullyFusedMLP flat params -> [(weight [out,in], bias [out])] in torch Linear layout.  tiny-cuda-nn networks have no biases; a
    tcnn.Network pads its input to 16 columns with ONES (Identity encoding, cpp_api.cu:151-153), so the padded columns of the first
    matrix act as a bias: bias_0 = sum of those columns (ones_padded_input=True).  All other biases are zero."""
    in_pad, out_pad = (in_dim + 15) // 16 * 16, (out_dim + 15) // 16 * 16
    need = tcnn_mlp_n_params(in_dim, width, n_hidden_layers, out_dim)
    if params.numel() < need:
        raise RuntimeError(f"tcnn MLP params: expected at least {need} values, got {params.numel()}")
    p = params.detach().float().reshape(-1)
    layers, off = [], 0
    w0 = p[off:off + width * in_pad].view(width, in_pad)
    off += width * in_pad
    b0 = w0[:, in_dim:].sum(1) if ones_padded_input and in_pad > in_dim else torch.zeros(width)
    layers.append((w0[:, :in_dim].clone(), b0.clone()))
    for _ in range(n_hidden_layers - 1):
        layers.append((p[off:off + width * width].view(width, width).clone(), torch.zeros(width)))
        off += width * width
    wl = p[off:off + out_pad * width].view(out_pad, width)
    layers.append((wl[:out_dim].clone(), torch.zeros(out_dim)))
    return layers


def unpack_tcnn_grid(params: torch.Tensor, n_levels: int, base_resolution: int, per_level_scale: float, log2_hashmap_size: int, features: int = 2):
    """GridEncoding flat params -> per level (rows [n_rows, F], level info)."""
    levels = tcnn_grid_levels(n_levels, base_resolution, per_level_scale, log2_hashmap_size)
    total = (levels[-1]["offset"] + levels[-1]["rows"]) * features
    if params.numel() != total:
        raise RuntimeError(f"tcnn grid params: expected {total} values for this configuration, got {params.numel()}")
    p = params.detach().float().reshape(-1, features)
    return [(p[lv["offset"]:lv["offset"] + lv["rows"]].clone(), lv) for lv in levels]


def convert_tcnn_state(model_state: Dict[str, torch.Tensor], config) -> Dict:
    """Splits every tinycudann `params` tensor of a NeRF-VO checkpoint (implementation='tcnn') into named pieces:
        {"field.mlp_base": {"mlp": [(W, b)...], "grid": [(rows, info)...]}, "field.mlp_head": {"mlp": [...]}, ...}
    MLP pieces are exact in torch Linear layout.  Grid levels keep tiny-cuda-nn's own row order: a HASHED level has the same vertex -> row
    hash as the torch path (primes 1, 2654435761, 805459861, grid.h / encodings.py:405-422) and can be copied row for row; a DENSE level
    (coarse resolutions) is indexed x + y*res + z*res^2 and has no counterpart in the always-hashed torch layout.  Either way tiny-cuda-nn
    interpolates at x*scale + 0.5 with scale = 2^(l*log2 s)*N - 1, not at x*floor(N*s^l), so the converted field is NOT the trained one:
    use this to inspect or to initialise, not to claim parity."""
    c = config
    growth = math.exp((math.log(c.max_res) - math.log(c.base_res)) / (c.num_levels - 1))
    out: Dict = {}

    def find(prefix):
        for k, v in model_state.items():
            if k.startswith(prefix) and k.endswith(".params"):
                return v
        return None

    p = find("field.mlp_base.")
    if p is not None:
        n_mlp = tcnn_mlp_n_params(c.num_levels * c.features_per_level, c.hidden_dim, 1, 16)
        out["field.mlp_base"] = {"mlp": unpack_tcnn_mlp(p[:n_mlp], c.num_levels * c.features_per_level, c.hidden_dim, 1, 16, ones_padded_input=False),
                                 "grid": unpack_tcnn_grid(p[n_mlp:], c.num_levels, c.base_res, growth, c.log2_hashmap_size, c.features_per_level)}
    p = find("field.mlp_head.")
    if p is not None:
        out["field.mlp_head"] = {"mlp": unpack_tcnn_mlp(p, 16 + 15 + c.appearance_embed_dim, c.hidden_dim_color, 2, 3)}
    p = find("field.mlp_pred_normals.")
    if p is not None:
        out["field.mlp_pred_normals"] = {"mlp": unpack_tcnn_mlp(p, 15 + 12, 64, 2, c.hidden_dim_transient)}
    for i, a in enumerate(c.proposal_net_args_list[:c.num_proposal_iterations]):
        g = math.exp((math.log(a["max_res"]) - math.log(16)) / (a["num_levels"] - 1))
        pe = find(f"proposal_networks.{i}.encoding.")
        if pe is None:
            pe = find(f"proposal_networks.{i}.mlp_base.0.")
        pm = find(f"proposal_networks.{i}.mlp_base.1.")
        entry = {}
        if pe is not None:
            entry["grid"] = unpack_tcnn_grid(pe, a["num_levels"], 16, g, a["log2_hashmap_size"], 2)
        if pm is not None:
            entry["mlp"] = unpack_tcnn_mlp(pm, a["num_levels"] * 2, a["hidden_dim"], 1, 1)
        if entry:
            out[f"proposal_networks.{i}"] = entry
    return out
