"""NerfactoModel / DepthNerfactoModel / ExtendedNerfactoModel — the model glue NeRF-VO's mapping thread trains
(NS/models/nerfacto.py:55-421, NS/models/depth_nerfacto.py:34-157, nerf_vo/mapping/nerfstudio_utils.py:326-350,
NS/models/base_model.py:131-192), wired to the nvo_b200 fields / samplers / renderers."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from . import losses as L
from .field_components import SceneContraction
from .fields import FieldHeadNames, HashMLPDensityField, NerfactoField
from .ray_samplers import ProposalNetworkSampler
from .rays import RayBundle
from .renderers import render_all


@dataclass
class NerfactoModelConfig:
    """Defaults = NS/models/nerfacto.py:55-131 with NeRF-VO's overrides (nerf_vo/mapping/nerfstudio.py:71-82)."""

    near_plane: float = 0.05
    far_plane: float = 1000.0
    hidden_dim: int = 64
    hidden_dim_color: int = 64
    hidden_dim_transient: int = 64
    num_levels: int = 16
    base_res: int = 16
    max_res: int = 2048
    log2_hashmap_size: int = 19
    features_per_level: int = 2
    num_proposal_samples_per_ray: Tuple[int, ...] = (256, 96)
    num_nerf_samples_per_ray: int = 48
    proposal_update_every: int = 5
    proposal_warmup: int = 5000
    num_proposal_iterations: int = 2
    proposal_net_args_list: List[Dict] = field(
        default_factory=lambda: [
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 128, "use_linear": False},
            {"hidden_dim": 16, "log2_hashmap_size": 17, "num_levels": 5, "max_res": 256, "use_linear": False},
        ]
    )
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    orientation_loss_mult: float = 0.0
    pred_normal_loss_mult: float = 0.0
    use_proposal_weight_anneal: bool = True
    use_average_appearance_embedding: bool = True
    proposal_weights_anneal_slope: float = 10.0
    proposal_weights_anneal_max_num_iters: int = 1000
    use_single_jitter: bool = True
    predict_normals: bool = True
    appearance_embed_dim: int = 32
    eval_num_rays_per_chunk: int = 4096
    # DepthNerfactoModelConfig (NS/models/depth_nerfacto.py:34-52) as NeRF-VO sets it
    depth_loss_mult: float = 0.001
    is_euclidean_depth: bool = False
    depth_sigma: float = 0.001
    # ExtendedNerfactoModelConfig (nerf_vo/mapping/nerfstudio_utils.py:326-330)
    normal_loss_mult: float = 0.000005
    # "fp16": field MLPs + hash features on the tcgen05 tensor-core path; "fp32": exact SIMT kernels (parity runs)
    precision: str = "fp16"
    # CameraOptimizerConfig.mode of the model's camera optimizer (NS/models/nerfacto.py:130 defaults to "SO3xR3", which is what NeRF-VO's
    # mapping step trains with, group "camera_opt", nerf_vo/mapping/nerfstudio.py:93-100).  "off" here: every golden vector of the parity
    # tests was generated with the optimizer off; MappingTrainer(camera_opt=True) and bench.py switch it on.
    camera_optimizer_mode: str = "off"


class NerfactoModel(nn.Module):
    config: NerfactoModelConfig

    def __init__(self, config: NerfactoModelConfig, num_train_data: int, aabb: Optional[torch.Tensor] = None) -> None:
        super().__init__()
        self.config = config
        self.num_train_data = num_train_data
        aabb = aabb if aabb is not None else torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
        contraction = SceneContraction(order=float("inf"))
        c = config
        self.field = NerfactoField(aabb, hidden_dim=c.hidden_dim, num_levels=c.num_levels, max_res=c.max_res, base_res=c.base_res,
                                   features_per_level=c.features_per_level, log2_hashmap_size=c.log2_hashmap_size, hidden_dim_color=c.hidden_dim_color,
                                   hidden_dim_transient=c.hidden_dim_transient, spatial_distortion=contraction, num_images=num_train_data,
                                   use_pred_normals=c.predict_normals, use_average_appearance_embedding=c.use_average_appearance_embedding,
                                   appearance_embedding_dim=c.appearance_embed_dim, precision=c.precision,
                                   pred_normals_trainable=c.pred_normal_loss_mult != 0)
        self.proposal_networks = nn.ModuleList()
        for i in range(c.num_proposal_iterations):
            args = c.proposal_net_args_list[min(i, len(c.proposal_net_args_list) - 1)]
            self.proposal_networks.append(HashMLPDensityField(aabb, spatial_distortion=contraction, **args))
        self.density_fns = [net.density_fn for net in self.proposal_networks]
        from .data import CameraOptimizerConfig

        self.camera_optimizer = CameraOptimizerConfig(mode=c.camera_optimizer_mode).setup(num_cameras=num_train_data, device="cpu")

        def update_schedule(step):
            return np.clip(np.interp(step, [0, c.proposal_warmup], [0, c.proposal_update_every]), 1, c.proposal_update_every)

        self.proposal_sampler = ProposalNetworkSampler(num_nerf_samples_per_ray=c.num_nerf_samples_per_ray,
                                                       num_proposal_samples_per_ray=c.num_proposal_samples_per_ray,
                                                       num_proposal_network_iterations=c.num_proposal_iterations, single_jitter=c.use_single_jitter,
                                                       update_sched=update_schedule)
        self.step = 0
        self._near_far_cache: Dict = {}

    # ---- bookkeeping the reference does through training callbacks (nerfacto.py:244-286) ------------------
    def get_param_groups(self) -> Dict[str, List[nn.Parameter]]:
        groups = {"proposal_networks": list(self.proposal_networks.parameters()), "fields": list(self.field.parameters())}
        self.camera_optimizer.get_param_groups(groups)
        return groups

    def before_train_iteration(self, step: int) -> None:
        self.step = step
        if self.config.use_proposal_weight_anneal:
            n = self.config.proposal_weights_anneal_max_num_iters
            frac = float(np.clip(step / n, 0, 1))
            b = self.config.proposal_weights_anneal_slope
            self.proposal_sampler.set_anneal(b * frac / ((b - 1) * frac + 1))

    def after_train_iteration(self, step: int) -> None:
        self.proposal_sampler.step_cb(step)

    # ---- forward -------------------------------------------------------------------------------------------------
    def set_nears_and_fars(self, ray_bundle: RayBundle) -> RayBundle:
        """NearFarCollider (NS/model_components/scene_colliders.py:186-191)."""
        near = self.config.near_plane if self.training else 0.0
        shape, dev = tuple(ray_bundle.origins[..., 0:1].shape), ray_bundle.origins.device
        key = (shape, str(dev), near, self.config.far_plane)
        if self._near_far_cache.get("key") != key:  # constants: built once per (batch shape, mode), not two kernels per step
            self._near_far_cache = {"key": key, "nears": torch.full(shape, near, dtype=torch.float32, device=dev),
                                    "fars": torch.full(shape, self.config.far_plane, dtype=torch.float32, device=dev)}
        ray_bundle.nears = self._near_far_cache["nears"]
        ray_bundle.fars = self._near_far_cache["fars"]
        return ray_bundle

    def forward(self, ray_bundle: RayBundle, jitters: Optional[List[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        return self.get_outputs(self.set_nears_and_fars(ray_bundle), jitters)

    def get_outputs(self, ray_bundle: RayBundle, jitters: Optional[List[torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        # apply the camera optimizer pose tweaks (NS/models/nerfacto.py:288-291) unless the step prologue already did
        if self.training and not ray_bundle.pose_corrected:
            self.camera_optimizer.apply_to_raybundle(ray_bundle)
        ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns, jitters=jitters)
        hook = getattr(self, "_before_field_forward", None)
        if hook is not None:
            hook()  # MappingTrainer(defer_fields_update=True): the previous step's fields update lands here, behind the proposal sampling
        fo = self.field.forward(ray_samples, compute_normals=self.config.predict_normals)
        hook = getattr(self, "_after_field_forward", None)
        if hook is not None:
            hook()  # MappingTrainer: the zero fill of the fields group's gradient range starts here, next to the render / loss kernels
        weights = ray_samples.get_weights(fo[FieldHeadNames.DENSITY])
        weights_list.append(weights)
        ray_samples_list.append(ray_samples)
        rgb, acc, dexp, dmed, _, n_img, pn_img = render_all(
            weights, ray_samples, rgb=fo[FieldHeadNames.RGB],
            normals=fo.get(FieldHeadNames.NORMALS), pred_normals=fo.get(FieldHeadNames.PRED_NORMALS), eval_mode=not self.training)
        outputs = {"rgb": rgb, "accumulation": acc, "depth": dmed, "expected_depth": dexp}
        if self.config.predict_normals:
            outputs["normals"] = n_img
            outputs["pred_normals"] = pn_img
        if self.training:
            outputs["weights_list"] = weights_list
            outputs["ray_samples_list"] = ray_samples_list
            if self.config.predict_normals:
                # multipliers are 0 in NeRF-VO (nerf_vo/mapping/nerfstudio.py:74-75): evaluated only when they can matter
                if self.config.orientation_loss_mult != 0:
                    outputs["rendered_orientation_loss"] = L.orientation_loss(weights.detach(), fo[FieldHeadNames.NORMALS], ray_bundle.directions)
                if self.config.pred_normal_loss_mult != 0:
                    outputs["rendered_pred_normal_loss"] = L.pred_normal_loss(weights.detach(), fo[FieldHeadNames.NORMALS].detach(),
                                                                            fo[FieldHeadNames.PRED_NORMALS])
        with torch.no_grad():
            from . import ops

            for i in range(self.config.num_proposal_iterations):
                if getattr(self, "_leaf_renders", False) and ops.leaf_streams.enabled:
                    # nothing inside the step consumes the proposal depth maps: under the trainer (which joins the side streams before
                    # the optimizer) they are rendered next to the loss kernels instead of in front of them
                    with ops.leaf_streams.fork(weights_list[i], ray_samples_list[i], aux=True):
                        outputs[f"prop_depth_{i}"] = render_all(weights_list[i].detach(), ray_samples_list[i])[3]
                else:
                    outputs[f"prop_depth_{i}"] = render_all(weights_list[i].detach(), ray_samples_list[i])[3]
        if ray_bundle.metadata is not None and "directions_norm" in ray_bundle.metadata:
            outputs["directions_norm"] = ray_bundle.metadata["directions_norm"]
        return outputs

    # ---- losses ----------------------------------------------------------------------------------------------------
    def get_metrics_dict(self, outputs, batch) -> Dict[str, torch.Tensor]:
        metrics = {}
        if self.training:
            metrics["distortion"] = L.distortion_loss(outputs["weights_list"], outputs["ray_samples_list"])
        return metrics

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        loss = {"rgb_loss": L.rgb_mse_loss(batch["image"], outputs["rgb"])}
        if self.training:
            c = self.config
            loss["interlevel_loss"] = c.interlevel_loss_mult * L.interlevel_loss(outputs["weights_list"], outputs["ray_samples_list"])
            assert metrics_dict is not None and "distortion" in metrics_dict
            loss["distortion_loss"] = c.distortion_loss_mult * metrics_dict["distortion"]
            if c.predict_normals:
                if "rendered_orientation_loss" in outputs:
                    loss["orientation_loss"] = c.orientation_loss_mult * torch.mean(outputs["rendered_orientation_loss"])
                if "rendered_pred_normal_loss" in outputs:
                    loss["pred_normal_loss"] = c.pred_normal_loss_mult * torch.mean(outputs["rendered_pred_normal_loss"])
            self.camera_optimizer.get_loss_dict(loss)  # nerfacto.py:379-380
        return loss

    # ---- evaluation (NS/models/base_model.py:164-192) ----------------------------------------------------------------
    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle: RayBundle, num_rays_per_chunk: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """Chunked evaluation of a whole-image bundle (NS/models/base_model.py:164-192).  Each output is allocated once at image size
        and every chunk writes its slice (the reference appends 200 chunk tensors per key and concatenates).  Rays are independent,
        so the result does not depend on the chunk size — except `expected_depth`, which the reference clips to the min / max sample
        distance of each CALL (renderers.py:379), i.e. per chunk there too.  evaluation/nerf_renderer.py renders with the config's
        4096-ray chunks; `NerfstudioRenderer` here passes a larger chunk (same pixels, fewer launches)."""
        chunk = num_rays_per_chunk or self.config.eval_num_rays_per_chunk
        image_shape = camera_ray_bundle.origins.shape[:-1]
        flat = camera_ray_bundle.flatten()
        n = len(flat)
        outs: Dict[str, torch.Tensor] = {}
        for i in range(0, n, chunk):
            o = self.forward(flat[i:i + chunk])
            for k, v in o.items():
                if not torch.is_tensor(v):
                    continue
                if k not in outs:
                    outs[k] = torch.empty((n,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
                outs[k][i:i + v.shape[0]] = v
        return {k: v.view(*image_shape, -1) for k, v in outs.items()}


class DepthNerfactoModel(NerfactoModel):
    """Adds the DS-NeRF depth loss over all three weight sets (NS/models/depth_nerfacto.py:79-125)."""

    def get_metrics_dict(self, outputs, batch):
        metrics = super().get_metrics_dict(outputs, batch)
        if self.training and "depth_image" in batch:
            n = len(outputs["weights_list"])
            d = 0.0
            for i in range(n):
                d = d + L.depth_loss(outputs["weights_list"][i], outputs["ray_samples_list"][i], batch["depth_image"], outputs["depth"],
                                     self.config.depth_sigma, outputs["directions_norm"], self.config.is_euclidean_depth) / n
            metrics["depth_loss"] = d
        return metrics

    def get_loss_dict(self, outputs, batch, metrics_dict=None):
        loss = super().get_loss_dict(outputs, batch, metrics_dict)
        if self.training and metrics_dict is not None and "depth_loss" in metrics_dict:
            loss["depth_loss"] = self.config.depth_loss_mult * metrics_dict["depth_loss"]
        return loss


class ExtendedNerfactoModel(DepthNerfactoModel):
    """Adds the MonoSDF normal loss on the rendered density-gradient normals (nerf_vo/mapping/nerfstudio_utils.py:333-350)."""

    def get_metrics_dict(self, outputs, batch):
        metrics = super().get_metrics_dict(outputs, batch)
        if "normal_image" in batch and self.config.normal_loss_mult > 0.0:
            metrics["normal_loss"] = L.monosdf_normal_loss(normal_pred=outputs["normals"], normal_gt=batch["normal_image"])
        return metrics

    def get_loss_dict(self, outputs, batch, metrics_dict=None):
        loss = super().get_loss_dict(outputs, batch, metrics_dict)
        if metrics_dict is not None and "normal_loss" in metrics_dict:
            loss["normal_loss"] = self.config.normal_loss_mult * metrics_dict["normal_loss"]
        return loss

    def get_train_loss_fused(self, ray_bundle: RayBundle, batch, jitters=None, eager_grads: bool = False):
        """Same step as get_train_loss_dict with every loss evaluated by ONE autograd node (ops.fused_step_losses): returns
        (outputs, total_loss, terms, weights): `total_loss` carries the graph; terms[name] * weights[name] are the entries
        get_loss_dict would produce (`terms` = ONE detached device vector viewed per name, so logging costs no kernels unless
        asked for).  Used by MappingTrainer; values and gradients equal the unfused path (tests/test_gpu_parity.py)."""
        from . import ops

        c = self.config
        assert self.training and c.num_proposal_iterations == 2, "fused losses cover the NeRF-VO training configuration"
        outputs = self(ray_bundle, jitters)
        wl, rl = outputs["weights_list"], outputs["ray_samples_list"]
        use_n = "normal_image" in batch and c.normal_loss_mult > 0.0 and c.predict_normals
        use_d = "depth_image" in batch
        mults = (1.0, c.interlevel_loss_mult, c.distortion_loss_mult, c.depth_loss_mult / len(wl) if use_d else 0.0, c.normal_loss_mult if use_n else 0.0)
        dnorm = outputs["directions_norm"] if not c.is_euclidean_depth else torch.ones_like(batch["depth_image"])
        total, terms = ops.fused_step_losses(
            wl, [r.sdist() for r in rl], [r.frustums.intervals() for r in rl], outputs["rgb"], batch["image"],
            normals_img=outputs["normals"] if use_n else None, normal_gt=batch["normal_image"] if use_n else None,
            depth_gt=batch["depth_image"] if use_d else None, directions_norm=dnorm if use_d else None, sigma=c.depth_sigma, mults=mults,
            eager_grads=eager_grads)
        if self.camera_optimizer.config.mode != "off" and ops.ray_grad_sink is None:
            reg: Dict[str, torch.Tensor] = {}
            self.camera_optimizer.get_loss_dict(reg)  # under MappingTrainer (ray_grad_sink set) the trainer adds value and gradient itself
            total = total + reg["camera_opt_regularizer"]
        names = ("rgb_loss", "interlevel_loss", "distortion_loss", "depth_loss", "normal_loss")
        # depth: `terms` holds the SUM over the weight sets, its weight the multiplier / number of sets (depth_nerfacto.py:93-103)
        return outputs, total, {n: terms[i] for i, n in enumerate(names) if mults[i] != 0.0}, {n: mults[i] for i, n in enumerate(names) if mults[i] != 0.0}

    def get_train_loss_dict(self, ray_bundle: RayBundle, batch, jitters=None):
        """VanillaPipeline.get_train_loss_dict (NS/pipelines/base_pipeline.py:291-304) for this model."""
        outputs = self(ray_bundle, jitters)
        metrics = self.get_metrics_dict(outputs, batch)
        return outputs, self.get_loss_dict(outputs, batch, metrics), metrics
