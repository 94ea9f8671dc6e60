"""Point-cloud export (meshing branch): the per-batch kernel against the oracle's restatement of
NS/exporter/exporter_utils.py:130-180,222-226, the host-side outlier filter and the PLY writer."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import nerfacto_oracle as O


def _batch(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g)
    d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    depth = 0.1 + 3.0 * torch.rand(n, 1, generator=g)
    acc = torch.rand(n, 1, generator=g)
    acc[::11] = 0.5  # exactly on the threshold: dropped (strict >)
    nc = torch.rand(n, 3, generator=g)
    return o, d, depth, acc, nc


def test_oracle_point_cloud_masks():
    o, d, depth, acc, nc = _batch(1000)
    p, idx, nrm = O.point_cloud_batch(o, d, depth, acc, nc, (-1.0, -1.0, -1.0), (1.0, 1.5, 2.0), True)
    assert p.shape[0] == idx.shape[0] == nrm.shape[0] and 0 < p.shape[0] < 1000
    assert bool((acc[idx, 0] > 0.5).all())
    assert bool((p > torch.tensor([-1.0, -1.0, -1.0])).all()) and bool((p < torch.tensor([1.0, 1.5, 2.0])).all())
    assert bool((torch.sum(d[idx] * nrm, -1) <= 0).all())  # re-oriented against the view direction


def test_statistical_outlier_filter_and_ply(tmp_path):
    import nerf_vo_b200 as nv
    from nerf_vo_b200 import exporter

    rng = np.random.default_rng(0)
    pts = rng.normal(size=(2000, 3)) * 0.1
    pts[:5] += 50.0  # five far outliers, far from each other too
    pts[:5] *= np.arange(1, 6)[:, None]
    ind = exporter.remove_statistical_outlier(pts, 20, 2.0)
    assert not set(range(5)) & set(ind.tolist()) and len(ind) > 1900
    pcd = exporter.PointCloud(pts[ind], rng.random((len(ind), 3)), rng.normal(size=(len(ind), 3)))
    f = tmp_path / "cloud.ply"
    exporter.write_ply(str(f), pcd)
    raw = open(f, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert f"element vertex {len(ind)}".encode() in head and len(body) == len(ind) * (6 * 4 + 3)
    rec = np.frombuffer(body, dtype=[("p", "<f4", 3), ("n", "<f4", 3), ("c", "u1", 3)])
    assert np.allclose(rec["p"], pcd.points.astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("box", [False, True])
@pytest.mark.parametrize("n", [1, 4097, 32768])
def test_point_cloud_kernel_vs_oracle(n, box):
    import nerf_vo_b200 as nv
    from nerf_vo_b200 import exporter

    o, d, depth, acc, nc = _batch(n, seed=n)
    lo, hi = ((-1.0, -1.2, -0.8), (1.1, 1.5, 2.0)) if box else (None, None)
    p_ref, idx_ref, n_ref = O.point_cloud_batch(o, d, depth, acc, nc, lo, hi, True)
    dev = "cuda:0"
    pts, nrm, keep = exporter.point_cloud_batch(o.to(dev), d.to(dev), depth.to(dev), acc.to(dev), nc.to(dev), lo, hi, True)
    assert torch.equal(torch.nonzero(keep)[:, 0].cpu(), idx_ref)  # the same rays survive, in the same order
    assert torch.equal(pts[keep].cpu(), p_ref)  # separately rounded mul / add: bit-identical to torch
    dot = torch.sum(d[idx_ref] * ((nc[idx_ref] * 2.0) - 1.0), -1)
    sure = dot.abs() > 1e-6  # the flip test compares a 3-term fp32 sum with 0: only ties may differ
    assert torch.equal(nrm[keep].cpu()[sure], n_ref[sure])


@pytest.mark.gpu
def test_generate_point_cloud_end_to_end():
    """generate_point_cloud over a small random-init model + synthetic keyframes, the way render_mesh drives it (`normals` output, re-oriented,
    axis-aligned box): the loop stops once num_points survivors are collected, every survivor passes the reference's two masks, normals face
    the camera side; a box that holds nothing ends through max_batches."""
    import nerf_vo_b200 as nv
    from nerf_vo_b200 import exporter
    from nerf_vo_b200.synthetic import synthetic_keyframes

    torch.manual_seed(0)
    dev = "cuda:0"
    K, H, W, B = 6, 48, 64, 2048
    dm = nv.DynamicDataManager(nv.DynamicDataManagerConfig(train_num_rays_per_batch=B, num_frames=K, frame_height=H, frame_width=W), device=dev)
    synthetic_keyframes(dm.train_dataset)
    dm.train_dataset.num_active_frames = K
    model = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(log2_hashmap_size=14), num_train_data=K).to(dev)
    lo, hi = (-3.0, -3.0, -3.0), (3.0, 3.0, 3.0)
    pcd = exporter.generate_point_cloud(model=model, datamanager=dm, num_points=3000, remove_outliers=False, reorient_normals=True,
                                        normal_output_name="normals", use_bounding_box=True, bounding_box_min=lo, bounding_box_max=hi, max_batches=60)
    assert len(pcd) >= 3000, len(pcd)  # random-init trunc_exp density ~ 1 over metres of ray: opaque
    assert len(pcd) < 3000 + B  # stopped after the batch that crossed num_points
    assert bool((pcd.points > np.array(lo)).all()) and bool((pcd.points < np.array(hi)).all())
    assert pcd.colors.min() >= 0 and pcd.colors.max() <= 1 and np.abs(pcd.normals).max() <= 1.0 + 1e-6
    none = exporter.generate_point_cloud(model=model, datamanager=dm, num_points=10, remove_outliers=False, use_bounding_box=True,
                                         bounding_box_min=(100, 100, 100), bounding_box_max=(101, 101, 101), max_batches=2)
    assert len(none) == 0 and none.normals is None
