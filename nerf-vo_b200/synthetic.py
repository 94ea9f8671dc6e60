"""Seeded synthetic Replica-shaped inputs for benchmarks and demos: pinhole rays from K random
cameras (fx=fy=600, cx=599.5, cy=339.5, 1200x680; /root/reference/datasets/replica.json:3-8), pixel sampling as
NS/data/pixel_samplers.py:103-106, directions as NS/cameras/cameras.py:620-654,875-878; rgb/depth/normal targets."""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F


def synthetic_rays(num_rays: int, num_images: int = 192, seed: int = 1234, H: int = 680, W: int = 1200, fx: float = 600.0, fy: float = 600.0,
                   cx: float = 599.5, cy: float = 339.5) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
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
    rays = {"origins": cam_o[c].contiguous(), "directions": (dworld / dnorm).contiguous(), "camera_indices": c[:, None].contiguous(),
            "directions_norm": dnorm.contiguous(), "pixel_area": torch.full((num_rays, 1), 1.0 / (fx * fy))}
    depth = torch.rand(num_rays, 1, generator=g) * 4.7 + 0.3
    depth = depth * (torch.rand(num_rays, 1, generator=g) > 0.1)
    targets = {"rgb": torch.rand(num_rays, 3, generator=g), "depth": depth, "normal": F.normalize(torch.randn(num_rays, 3, generator=g), dim=-1)}
    return rays, targets


def synthetic_jitters(num_rays: int, n_levels: int = 3, seed: int = 99) -> List[torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(num_rays, 1, generator=g) for _ in range(n_levels)]


def synthetic_keyframes(dataset, seed: int = 4321, fx: float = 320.0, fy: float = 320.0) -> None:
    """Fills a DynamicDataset (data.py) in place with seeded keyframes: Replica intrinsics scaled to the frame size
    (/root/reference/configs/nerf_vo_replica.yaml:16-17), random poses (origins U[-0.5,0.5]^3), uniform colour, depth U[0.3,5] with
    10 % invalid (0), random unit camera-frame normals.  Generated on the dataset's device (1.2 GB at 192 x 360 x 640)."""
    K, H, W, dev = dataset.num_frames, dataset.frame_height, dataset.frame_width, dataset.device
    g = torch.Generator(device=dev).manual_seed(seed)
    q = F.normalize(torch.randn(K, 4, generator=g, device=dev), dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1).reshape(K, 3, 3)
    dataset.camera_intrinsics[:] = torch.tensor([fx, fy, W / 2 - 0.5, H / 2 - 0.5], device=dev)
    dataset.camera_extrinsics[:, :3, :3] = R
    dataset.camera_extrinsics[:, :3, 3] = torch.rand(K, 3, generator=g, device=dev) - 0.5
    dataset.frames_color.uniform_(generator=g)
    dataset.frames_depth.uniform_(generator=g).mul_(4.7).add_(0.3)
    dataset.frames_depth.mul_((torch.rand(dataset.frames_depth.shape, generator=g, device=dev) > 0.1).float())
    if dataset.frames_normal is not None:
        dataset.frames_normal.normal_(generator=g)
        dataset.frames_normal.div_(dataset.frames_normal.norm(dim=-1, keepdim=True).clamp_min(1e-12))
    dataset.num_active_frames = K
