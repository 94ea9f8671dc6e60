"""Point-cloud export (the meshing branch of the evaluation loop): `generate_point_cloud` with the reference's arguments
(NS/exporter/exporter_utils.py:78-231) as `NerfstudioRenderer.render_mesh` calls it for a `predict_normals` model
(evaluation/nerf_renderer.py:188-203: 33 554 432 points, 32 768 rays per batch, `normals` output, re-oriented normals, bounding box).

What runs where:
  * ray batches come from the datamanager's device-side step prologue (`next_train`), the model evaluates them through the same kernels as
    the frame loop, and ONE kernel (csrc/frame.cu:k_point_cloud) turns a batch into surface points, the keep mask of the reference's opacity
    and box tests, and decoded, view-re-oriented normals; the survivors are compacted on the device and cross PCIe once at the end;
  * `remove_statistical_outlier` (open3d in the reference: mean distance to the 20 nearest neighbours against mean + std_ratio * std) is a
    host post-process there and here (scipy's cKDTree); it is off the GPU path and optional;
  * Poisson reconstruction itself (open3d `create_from_point_cloud_poisson`) is outside the scope of this library: `write_ply` hands the
    oriented point cloud to whatever mesher the caller uses.
The reference re-orients after outlier removal on the host; the per-point test is the same either way, so it is applied in the kernel."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import call, check


@dataclass
class PointCloud:
    points: np.ndarray   # [N,3] float64 (open3d Vector3dVector of the reference)
    colors: np.ndarray   # [N,3] float64 in [0,1]
    normals: Optional[np.ndarray]  # [N,3] float64 or None

    def __len__(self) -> int:
        return int(self.points.shape[0])


def point_cloud_batch(origins, directions, depth, accumulation, normals_coded=None, bounding_box_min=None, bounding_box_max=None, reorient_normals=False):
    """One ray batch -> (points [n,3], normals [n,3] | None, keep [n] bool): the per-ray part of generate_point_cloud's loop body."""
    n = origins.shape[0]
    origins = check(origins.reshape(n, 3).contiguous(), "origins", torch.float32, (n, 3))
    directions = check(directions.reshape(n, 3).contiguous(), "directions", torch.float32, (n, 3))
    depth = check(depth.reshape(n).contiguous(), "depth", torch.float32, (n,))
    accumulation = check(accumulation.reshape(n).contiguous(), "accumulation", torch.float32, (n,))
    if normals_coded is not None:
        normals_coded = check(normals_coded.reshape(n, 3).contiguous(), "normals", torch.float32, (n, 3))
    if (bounding_box_min is None) != (bounding_box_max is None):
        raise ValueError("bounding_box_min and bounding_box_max go together")
    lo = hi = None
    if bounding_box_min is not None:
        lo, hi = (ctypes.c_float * 3)(*[float(v) for v in bounding_box_min]), (ctypes.c_float * 3)(*[float(v) for v in bounding_box_max])
        if not all(a < b for a, b in zip(lo, hi)):
            raise AssertionError(f"Bounding box min {tuple(bounding_box_min)} must be smaller than max {tuple(bounding_box_max)}")  # exporter_utils.py:160-162
    points = torch.empty((n, 3), dtype=torch.float32, device=origins.device)
    normals = torch.empty((n, 3), dtype=torch.float32, device=origins.device) if normals_coded is not None else None
    keep = torch.empty((n,), dtype=torch.uint8, device=origins.device)
    call("nvo_point_cloud", n, origins, directions, depth, accumulation, normals_coded, None if lo is None else ctypes.addressof(lo),
         None if hi is None else ctypes.addressof(hi), int(bool(reorient_normals)), points, normals, keep)
    return points, normals, keep.bool()


def remove_statistical_outlier(points: np.ndarray, nb_neighbors: int = 20, std_ratio: float = 10.0) -> np.ndarray:
    """Indices kept by open3d's PointCloud.remove_statistical_outlier: a point stays while the mean distance to its nb_neighbors nearest
    neighbours (itself included, as open3d's KNN search returns it) is below mean + std_ratio * std of that statistic over the cloud."""
    from scipy.spatial import cKDTree

    if points.shape[0] <= nb_neighbors:
        return np.arange(points.shape[0])
    d, _ = cKDTree(points).query(points, k=nb_neighbors)
    avg = d.mean(axis=1)
    thr = avg.mean() + std_ratio * avg.std()
    return np.nonzero(avg < thr)[0]


@torch.no_grad()
def generate_point_cloud(pipeline=None, num_points: int = 1000000, remove_outliers: bool = True, estimate_normals: bool = False,
                         reorient_normals: bool = False, rgb_output_name: str = "rgb", depth_output_name: str = "depth",
                         normal_output_name: Optional[str] = None, use_bounding_box: bool = True,
                         bounding_box_min: Optional[Tuple[float, float, float]] = None, bounding_box_max: Optional[Tuple[float, float, float]] = None,
                         crop_obb=None, std_ratio: float = 10.0, model=None, datamanager=None, max_batches: Optional[int] = None) -> PointCloud:
    """exporter_utils.generate_point_cloud over an nvo_b200 model: `pipeline` (anything with .model and .datamanager) or model= / datamanager=.
    Batches are drawn with datamanager.next_train(0) until num_points survivors are collected (`max_batches` bounds the loop for scenes whose
    rays are mostly transparent — the reference would spin forever there)."""
    if pipeline is not None:
        model, datamanager = pipeline.model, pipeline.datamanager
    if model is None or datamanager is None:
        raise ValueError("generate_point_cloud needs a pipeline or model= and datamanager=")
    if crop_obb is not None:
        raise NotImplementedError("oriented crop boxes are not on NeRF-VO's path (evaluation/nerf_renderer.py:192-203 passes an axis-aligned box)")
    if estimate_normals:
        if normal_output_name is not None:
            raise ValueError("Cannot estimate normals and use normal_output_name at the same time")  # exporter_utils.py:205-208
        raise NotImplementedError("estimate_normals (open3d's PCA normals) is not on NeRF-VO's path: it exports the model's `normals` output")
    box = use_bounding_box and bounding_box_min is not None
    was_training = model.training
    model.eval()
    pts, cols, nrm, got, batches = [], [], [], 0, 0
    try:
        while got < num_points and (max_batches is None or batches < max_batches):
            ray_bundle, _ = datamanager.next_train(0)
            outputs = model(ray_bundle)
            for name in (rgb_output_name, depth_output_name) + ((normal_output_name,) if normal_output_name is not None else ()):
                if name not in outputs:
                    raise KeyError(f"Could not find {name} in the model outputs; available: {sorted(outputs.keys())}")
            n_coded = outputs[normal_output_name] if normal_output_name is not None else None
            if n_coded is not None and not (float(n_coded.min()) >= 0.0 and float(n_coded.max()) <= 1.0):
                raise AssertionError("Normal values from method output must be in [0, 1]")  # exporter_utils.py:141-143
            points, normals, keep = point_cloud_batch(ray_bundle.origins, ray_bundle.directions, outputs[depth_output_name], outputs["accumulation"], n_coded,
                                                      bounding_box_min if box else None, bounding_box_max if box else None, reorient_normals)
            pts.append(points[keep])
            cols.append(outputs[rgb_output_name].reshape(-1, 3)[keep])
            if normals is not None:
                nrm.append(normals[keep])
            got += int(pts[-1].shape[0])
            batches += 1
    finally:
        model.train(was_training)
    points = torch.cat(pts).double().cpu().numpy() if pts else np.zeros((0, 3))
    colors = torch.cat(cols).double().cpu().numpy() if cols else np.zeros((0, 3))
    normals = torch.cat(nrm).double().cpu().numpy() if nrm else None
    if remove_outliers and points.shape[0] > 0:
        ind = remove_statistical_outlier(points, 20, std_ratio)
        points, colors = points[ind], colors[ind]
        if normals is not None:
            normals = normals[ind]
    return PointCloud(points, colors, normals)


def write_ply(path: str, pcd: PointCloud) -> None:
    """Binary little-endian PLY (x y z [nx ny nz] red green blue) — the file open3d's read_point_cloud and Poisson meshers take."""
    n = len(pcd)
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
    if pcd.normals is not None:
        fields += [("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")]
    fields += [("red", "u1"), ("green", "u1"), ("blue", "u1")]
    rec = np.zeros(n, dtype=fields)
    rec["x"], rec["y"], rec["z"] = pcd.points[:, 0], pcd.points[:, 1], pcd.points[:, 2]
    if pcd.normals is not None:
        rec["nx"], rec["ny"], rec["nz"] = pcd.normals[:, 0], pcd.normals[:, 1], pcd.normals[:, 2]
    c = np.clip(np.rint(pcd.colors * 255.0), 0, 255).astype(np.uint8)
    rec["red"], rec["green"], rec["blue"] = c[:, 0], c[:, 1], c[:, 2]
    names = {"<f4": "float", "u1": "uchar"}
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n" + "".join(f"property {names[t]} {k}\n" for k, t in fields) + "end_header\n"
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())
