"""Step prologue, host side: the reference's data-manager surface over the fused CUDA prologue.

Mirrors, with the same names / argument meaning:
  * `DynamicDataset`, `DynamicDataManager(.next_train)`  nerf_vo/mapping/nerfstudio_utils.py:30-241,243-305
  * `PixelSampler(.sample)`                              NS/data/pixel_samplers.py:60-219,300-317
  * `RayGenerator`, `Cameras(.generate_rays)`            NS/model_components/ray_generators.py:25-57, NS/cameras/cameras.py:503-912
  * `CameraOptimizer(.forward / .apply_to_raybundle)`    NS/cameras/camera_optimizers.py:60-147

`DynamicDataManager.next_train` is ONE kernel launch (csrc/batch.cu): random draws -> (camera,row,col), colour / depth / normal
gather at those pixels (normals rotated into the world frame per sampled pixel instead of re-solving every frame each step),
pinhole rays, pose correction.  The reference's `c.cpu(), y.cpu(), x.cpu()` sync (pixel_samplers.py:211) does not exist here:
everything stays on the device, so the prologue is CUDA-graph capturable.  No CPU fallback."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import ops
from .rays import RayBundle

POSE_MODES = {"off": 0, "SO3xR3": 1, "SE3": 2}


class Cameras:
    """Perspective cameras (NS/cameras/cameras.py:80-190): fx/fy/cx/cy [K], camera_to_worlds [K,3|4,4], one image size."""

    def __init__(self, fx, fy, cx, cy, height: int, width: int, camera_to_worlds: torch.Tensor, intrinsics: Optional[torch.Tensor] = None,
                 extrinsics: Optional[torch.Tensor] = None):
        self.height, self.width = int(height), int(width)
        dev = camera_to_worlds.device
        # packed views the kernels read: [K,4] fx fy cx cy, [K,4,4] c2w
        self.intrinsics = intrinsics if intrinsics is not None else torch.stack(
            [torch.as_tensor(v, dtype=torch.float32, device=dev).reshape(-1) for v in (fx, fy, cx, cy)], dim=-1).contiguous()
        if extrinsics is not None:
            self.extrinsics = extrinsics
        else:
            K = camera_to_worlds.shape[0]
            self.extrinsics = torch.eye(4, dtype=torch.float32, device=dev).repeat(K, 1, 1)
            self.extrinsics[:, :camera_to_worlds.shape[1]] = camera_to_worlds
        self.extrinsics = self.extrinsics.contiguous()

    fx = property(lambda self: self.intrinsics[:, 0:1])
    fy = property(lambda self: self.intrinsics[:, 1:2])
    cx = property(lambda self: self.intrinsics[:, 2:3])
    cy = property(lambda self: self.intrinsics[:, 3:4])
    camera_to_worlds = property(lambda self: self.extrinsics[:, :3])
    device = property(lambda self: self.extrinsics.device)

    def __len__(self) -> int:
        return self.extrinsics.shape[0]

    def to(self, device) -> "Cameras":
        if torch.device(device) == self.extrinsics.device:
            return self  # keeps aliasing the dataset's live buffers, as `.to` on an already-resident tensor does
        return Cameras(None, None, None, None, self.height, self.width, self.extrinsics[:, :3], self.intrinsics.to(device), self.extrinsics.to(device))

    def get_image_coords(self, pixel_offset: float = 0.5) -> torch.Tensor:
        """[H,W,2] (y,x) pixel centres (cameras.py:293-318)."""
        ys, xs = torch.meshgrid(torch.arange(self.height, device=self.device), torch.arange(self.width, device=self.device), indexing="ij")
        return torch.stack([ys, xs], dim=-1) + pixel_offset

    def generate_rays(self, camera_indices, coords: Optional[torch.Tensor] = None, keep_shape: Optional[bool] = None) -> RayBundle:
        """camera_indices: int -> every pixel of that frame as an [H,W] bundle (cameras.py:330-420 with coords=None);
        tensor [n,1] with coords [n,2] (y,x pixel centres, i.e. integer + 0.5) -> those rays (RayGenerator's call)."""
        if isinstance(camera_indices, int):
            if coords is not None:
                raise RuntimeError("generate_rays(int camera index) renders the full frame; pass a tensor of indices with coords")
            o, d, dn, pa, ci = ops.generate_rays(self.intrinsics, self.extrinsics, cam=camera_indices, height=self.height, width=self.width)
            H, W = self.height, self.width
            return RayBundle(origins=o.view(H, W, 3), directions=d.view(H, W, 3), pixel_area=pa.view(H, W, 1), camera_indices=ci.view(H, W, 1),
                             metadata={"directions_norm": dn.view(H, W, 1)})
        if coords is None:
            raise RuntimeError("generate_rays(tensor camera_indices) needs coords")
        idx = torch.cat([camera_indices.reshape(-1, 1).long(), torch.floor(coords.reshape(-1, 2)).long()], dim=-1).contiguous()
        o, d, dn, pa, ci = ops.generate_rays(self.intrinsics, self.extrinsics, indices=idx)
        return RayBundle(origins=o, directions=d, pixel_area=pa, camera_indices=ci, metadata={"directions_norm": dn})


class RayGenerator(nn.Module):
    """ray_indices [n,3] (camera,row,col) -> RayBundle (NS/model_components/ray_generators.py:25-57)."""

    def __init__(self, cameras: Cameras) -> None:
        super().__init__()
        self.cameras = cameras

    def forward(self, ray_indices: torch.Tensor) -> RayBundle:
        o, d, dn, pa, ci = ops.generate_rays(self.cameras.intrinsics, self.cameras.extrinsics, indices=ray_indices.long().contiguous())
        return RayBundle(origins=o, directions=d, pixel_area=pa, camera_indices=ci, metadata={"directions_norm": dn})


@dataclass
class PixelSamplerConfig:
    num_rays_per_batch: int = 4096

    def setup(self, **kw) -> "PixelSampler":
        return PixelSampler(self, **kw)


class PixelSampler:
    """Uniform pixel sampler (NS/data/pixel_samplers.py:60-219).  `sample(image_batch)` draws torch.rand((B,3)) like the reference
    (same generator stream => same pixels) and returns the collated pixel batch; masks / equirectangular / fisheye crops are
    not on NeRF-VO's path and raise."""

    def __init__(self, config: PixelSamplerConfig, **kw) -> None:
        self.config = config
        self.num_rays_per_batch = kw.get("num_rays_per_batch", config.num_rays_per_batch)

    def set_num_rays_per_batch(self, n: int) -> None:
        self.num_rays_per_batch = n

    def sample_method(self, batch_size: int, num_images: int, image_height: int, image_width: int, mask=None, device="cpu") -> torch.Tensor:
        if mask is not None:
            raise NotImplementedError("masked pixel sampling is not on the NeRF-VO mapping path")
        u = torch.rand((batch_size, 3), device=device)
        return (u * torch.tensor([num_images, image_height, image_width], device=device)).long()

    def sample(self, image_batch: Dict) -> Dict:
        if "mask" in image_batch:
            raise NotImplementedError("masked pixel sampling is not on the NeRF-VO mapping path")
        img = image_batch["image"]
        K, H, W, _ = img.shape
        indices = self.sample_method(self.num_rays_per_batch, K, H, W, device=img.device)
        c, y, x = indices.unbind(-1)
        out = {k: v[c, y, x] for k, v in image_batch.items() if k != "image_idx" and v is not None}
        indices[:, 0] = image_batch["image_idx"][c]
        out["indices"] = indices
        return out


@dataclass
class CameraOptimizerConfig:
    """NS/cameras/camera_optimizers.py:39-56."""

    mode: str = "off"
    trans_l2_penalty: float = 1e-2
    rot_l2_penalty: float = 1e-3

    def setup(self, num_cameras: int, device) -> "CameraOptimizer":
        return CameraOptimizer(self, num_cameras, device)


class CameraOptimizer(nn.Module):
    """Learnable per-camera pose deltas (NS/cameras/camera_optimizers.py:60-147).  NeRF-VO runs mode 'SE3'
    (nerf_vo/mapping/nerfstudio.py:64), the nerfstudio default is 'SO3xR3' (NS/models/nerfacto.py:130)."""

    def __init__(self, config: CameraOptimizerConfig, num_cameras: int, device, non_trainable_camera_indices=None) -> None:
        super().__init__()
        if config.mode not in POSE_MODES:
            raise ValueError(f"camera optimizer mode must be one of {list(POSE_MODES)}, got {config.mode!r}")
        self.config, self.num_cameras, self.device = config, num_cameras, device
        self.non_trainable_camera_indices = non_trainable_camera_indices
        if config.mode != "off":
            self.pose_adjustment = nn.Parameter(torch.zeros((num_cameras, 6), device=device))

    @property
    def mode_id(self) -> int:
        return POSE_MODES[self.config.mode]

    def forward(self, indices: torch.Tensor) -> torch.Tensor:
        """[n,3,4] correction matrices (identity when off)."""
        if self.config.mode == "off":
            return torch.eye(4, device=self.device)[None, :3, :4].tile(indices.shape[0], 1, 1)
        out = ops.pose_exp_map(self.pose_adjustment[indices, :].contiguous(), self.mode_id)
        if self.non_trainable_camera_indices is not None:
            out[self.non_trainable_camera_indices.to(out.device)] = torch.eye(4, device=out.device)[:3, :4]
        return out

    def apply_to_raybundle(self, raybundle: RayBundle) -> None:
        if self.config.mode == "off":
            return
        o, d = ops.pose_correction(raybundle.origins.contiguous(), raybundle.directions.contiguous(), raybundle.camera_indices.reshape(-1).contiguous(),
                                   self.pose_adjustment, self.mode_id)
        raybundle.origins, raybundle.directions = o, d

    def get_loss_dict(self, loss_dict: dict) -> None:
        """Regulariser on the pose deltas (camera_optimizers.py:149-155)."""
        if self.config.mode != "off":
            loss_dict["camera_opt_regularizer"] = (self.pose_adjustment[:, :3].norm(dim=-1).mean() * self.config.trans_l2_penalty
                                                   + self.pose_adjustment[:, 3:].norm(dim=-1).mean() * self.config.rot_l2_penalty)

    def get_metrics_dict(self, metrics_dict: dict) -> None:
        if self.config.mode != "off":
            metrics_dict["camera_opt_translation"] = self.pose_adjustment[:, :3].norm()
            metrics_dict["camera_opt_rotation"] = self.pose_adjustment[:, 3:].norm()

    def get_param_groups(self, param_groups: dict) -> None:
        ps = list(self.parameters())
        if self.config.mode != "off":
            assert len(ps) > 0
            param_groups["camera_opt"] = ps


class DynamicDataset(torch.utils.data.Dataset):
    """Keyframe store of the mapping thread (nerf_vo/mapping/nerfstudio_utils.py:30-241): preallocated device buffers the SLAM
    front-end fills; `num_active_frames` grows as keyframes arrive."""

    def __init__(self, num_frames: int, frame_height: int, frame_width: int, device=torch.device("cuda:0"), use_normals: bool = True) -> None:
        super().__init__()
        self.device, self.use_normals = torch.device(device), use_normals
        # device copy of `num_active_frames`: the step prologue reads it, so a captured CUDA graph samples the keyframes inserted
        # AFTER its capture as well (a by-value kernel argument would freeze the count at capture time)
        self.num_active_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.num_frames, self.num_active_frames = num_frames, 0
        self.frame_height, self.frame_width = frame_height, frame_width
        self.normalization_matrix = None
        f = dict(dtype=torch.float32, device=self.device)
        self.camera_intrinsics = torch.zeros((num_frames, 4), **f)
        self.camera_extrinsics = torch.eye(4, **f).repeat(num_frames, 1, 1)
        self.frames_color = torch.zeros((num_frames, frame_height, frame_width, 3), **f)
        self.frames_depth = torch.zeros((num_frames, frame_height, frame_width, 1), **f)
        self.frames_normal = torch.zeros((num_frames, frame_height, frame_width, 3), **f) if use_normals else None
        self.cameras = Cameras(None, None, None, None, frame_height, frame_width, self.camera_extrinsics[:, :3], self.camera_intrinsics, self.camera_extrinsics)

    @property
    def num_active_frames(self) -> int:
        return self._num_active_frames

    @num_active_frames.setter
    def num_active_frames(self, n: int) -> None:
        self._num_active_frames = int(n)
        self.num_active_dev.fill_(int(n))

    def __len__(self) -> int:
        return self.num_active_frames if self.num_active_frames > 0 else self.num_frames

    def update(self, input: dict) -> None:
        self.insert_update(self.prepare_update(input))

    def prepare_update(self, input: dict) -> dict:
        """nerfstudio_utils.py:160-212: channel-last frames, scene normalisation so that frame 0 is canonical."""
        assert int(input["keyframe_indices"].max()) < self.num_frames
        n_new = input["frames_color"].shape[0]
        if input["camera_extrinsics"].shape[0] == n_new:
            indices = input["keyframe_indices"]
            num_active = int(input["keyframe_indices"].max()) + 1
        else:
            indices = torch.arange(self.num_active_frames, self.num_active_frames + n_new)
            num_active = self.num_active_frames + n_new
        ext = input["camera_extrinsics"].detach().clone()
        if self.normalization_matrix is None:
            target = torch.tensor([[1, 0, 0, 0], [0, 0, -1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=ext.dtype, device=ext.device)
            self.normalization_matrix = torch.linalg.solve(ext[0], target, left=True)
        ext = (self.normalization_matrix @ ext.permute(2, 1, 0)).permute(2, 1, 0)
        out = {"indices": indices, "keyframe_indices": input["keyframe_indices"], "num_active_frames": num_active,
               "camera_intrinsics": input["camera_intrinsics"].detach().clone(), "camera_extrinsics": ext,
               "frames_color": input["frames_color"].detach().clone().permute(0, 2, 3, 1), "frames_depth": input["frames_depth"].detach().clone().permute(0, 2, 3, 1)}
        if self.use_normals:
            out["frames_normal"] = input["frames_normal"].detach().clone().permute(0, 2, 3, 1)
        return out

    def insert_update(self, input: dict) -> None:
        dev = self.device
        self.camera_intrinsics[input["indices"]] = input["camera_intrinsics"].to(dev)
        self.camera_extrinsics[input["keyframe_indices"].to(dev)] = input["camera_extrinsics"].to(dev)
        self.frames_color[input["indices"]] = input["frames_color"].to(dev)
        self.frames_depth[input["keyframe_indices"]] = input["frames_depth"].to(dev)
        if self.use_normals:
            self.frames_normal[input["indices"]] = input["frames_normal"].to(dev)
        self.num_active_frames = input["num_active_frames"]

    def get_dataset(self) -> dict:
        """The whole-frame dict of the reference (:133-155) — kept for API parity (viewer / export); the training path does not
        materialise it (next_train gathers the sampled pixels only)."""
        n = self.num_active_frames
        data = {"image_idx": torch.arange(0, n, dtype=torch.long, device=self.device), "image": self.frames_color[:n], "depth_image": self.frames_depth[:n]}
        if self.use_normals:
            H, W = self.frame_height, self.frame_width
            nrm = torch.linalg.solve(self.camera_extrinsics[:n, :3, :3], self.frames_normal[:n].permute(0, 3, 1, 2).reshape(n, 3, H * W))
            data["normal_image"] = (nrm.reshape(n, 3, H, W).permute(0, 2, 3, 1) + 1) / 2
        return data


@dataclass
class DynamicDataManagerConfig:
    train_num_rays_per_batch: int = 4096
    eval_num_rays_per_batch: int = 4096
    camera_optimizer: CameraOptimizerConfig = field(default_factory=CameraOptimizerConfig)
    num_frames: int = 128
    frame_height: int = 480
    frame_width: int = 640
    use_normals: bool = True


class DynamicDataManager:
    """nerf_vo/mapping/nerfstudio_utils.py:257-305.  `next_train(step)` -> (RayBundle, batch): one fused launch."""

    def __init__(self, config: DynamicDataManagerConfig, device=torch.device("cuda:0")):
        self.config, self.device = config, torch.device(device)
        self.train_dataset = DynamicDataset(config.num_frames, config.frame_height, config.frame_width, self.device, config.use_normals)
        self.eval_dataset = None
        self.train_pixel_sampler = PixelSamplerConfig().setup(num_rays_per_batch=config.train_num_rays_per_batch)
        self.train_ray_generator = RayGenerator(self.train_dataset.cameras.to(self.device))
        self.camera_optimizer: Optional[CameraOptimizer] = None  # NerfactoModel owns it in the reference (nerfacto.py:171); attach to fuse
        self._u: Optional[torch.Tensor] = None
        self._last_camera_indices: Optional[torch.Tensor] = None

    def get_train_rays_per_batch(self) -> int:
        return self.config.train_num_rays_per_batch

    def next_train(self, step: int, u: Optional[torch.Tensor] = None) -> Tuple[RayBundle, Dict]:
        """u: optional [B,3] uniform draws (tests / replay); default = torch.rand on the device, the reference's draw."""
        ds = self.train_dataset
        B = self.train_pixel_sampler.num_rays_per_batch
        if ds.num_active_frames <= 0:
            raise RuntimeError("next_train: the dataset holds no active frames")
        if u is None:
            if self._u is None or self._u.shape[0] != B:
                self._u = torch.empty((B, 3), dtype=torch.float32, device=self.device)
            u = self._u.uniform_()  # in place: the buffer address is stable across CUDA-graph replays
        co = self.camera_optimizer
        pose = co.pose_adjustment if co is not None and co.config.mode != "off" else None
        out = ops.batch_prologue(u, ds.num_active_frames, ds.camera_intrinsics, ds.camera_extrinsics, ds.frames_color, ds.frames_depth,
                                 ds.frames_normal if ds.use_normals else None, pose, co.mode_id if pose is not None else 0,
                                 num_active_dev=ds.num_active_dev, want_raw_directions=pose is not None)
        rb = RayBundle(origins=out["origins"], directions=out["directions"], pixel_area=out["pixel_area"], camera_indices=out["camera_indices"],
                       metadata={"directions_norm": out["directions_norm"]}, pose_corrected=pose is not None)
        self._last_directions_raw = out.get("directions_raw")  # saved input of the pose correction's backward
        self._last_camera_indices = out["camera_indices"]  # under a CUDA graph: refreshed by every replay (diagnostics / tests)
        batch = {"indices": out["indices"], "image": out["image"], "depth_image": out["depth_image"]}
        if ds.use_normals:
            batch["normal_image"] = out["normal_image"]
        return rb, batch

    def next_eval(self, step: int) -> Tuple[RayBundle, Dict]:
        return self.next_train(step)
