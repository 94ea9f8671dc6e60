"""Step prologue (pixel sampling + gather + ray generation + camera-pose correction).

CPU: the oracle restatement against the golden vectors made by the reference's own PixelSampler / RayGenerator / Cameras /
CameraOptimizer (oracle/make_golden.py:gen_batch_prologue), and the SE3 closed form against torch.linalg.matrix_exp
(lietorch, which the reference calls for SE3, is not installed anywhere here: that mode's parity is unpinned).
GPU: csrc/batch.cu through the C ABI against the golden vectors and the oracle.  Tolerances:
  * pixel / camera indices: bit-exact;  gathered colour / depth: bit-exact (pure copies);
  * rays: 2e-6 absolute on unit vectors / origins (fp32, different summation order than ATen), pixel_area 1e-6 relative;
  * normals (3x3 solve by adjugate vs LU): 1e-5;  pose matrices 1e-6;  pose gradients 1e-4 of max-abs.
"""
import numpy as np
import pytest
import torch

import nerfacto_oracle as O


def T(a, dev="cpu"):
    return torch.from_numpy(np.asarray(a)).to(dev)


# ---- CPU: oracle vs reference golden --------------------------------------------------------------------------------------
def test_oracle_prologue_matches_reference_golden(golden):
    g = {k: T(v) for k, v in golden("batch_prologue").items()}
    rays, batch = O.next_train_batch(g["u"], g["intrinsics"], g["extrinsics"], g["frames_color"], g["frames_depth"], g["frames_normal"])
    assert torch.equal(batch["indices"], g["indices"]) and torch.equal(rays["camera_indices"], g["camera_indices"])
    assert torch.equal(batch["image"], g["image"]) and torch.equal(batch["depth_image"], g["depth_image"])
    torch.testing.assert_close(batch["normal_image"], g["normal_image"], rtol=0, atol=1e-6)
    for k in ("origins", "directions", "pixel_area", "directions_norm"):
        torch.testing.assert_close(rays[k], g[k], rtol=1e-6, atol=1e-7, msg=k)


def test_oracle_pose_correction_matches_reference_golden(golden):
    g = {k: T(v) for k, v in golden("batch_prologue").items()}
    torch.testing.assert_close(O.exp_map_so3xr3(g["pose_adjustment"]), g["pose_matrices_so3xr3"], rtol=0, atol=1e-6)
    adj = g["pose_adjustment"].clone().requires_grad_(True)
    rays, _ = O.next_train_batch(g["u"], g["intrinsics"], g["extrinsics"], g["frames_color"], g["frames_depth"], None, adj, "SO3xR3")
    torch.testing.assert_close(rays["origins"].detach(), g["origins_so3xr3"], rtol=0, atol=1e-6)
    torch.testing.assert_close(rays["directions"].detach(), g["directions_so3xr3"], rtol=0, atol=1e-6)
    ((rays["origins"] * g["cot_origins"]).sum() + (rays["directions"] * g["cot_directions"]).sum()).backward()
    torch.testing.assert_close(adj.grad, g["pose_adjustment_grad"], rtol=1e-5, atol=1e-5)


def _se3_tangents():
    gen = torch.Generator().manual_seed(3)
    t = torch.randn(40, 6, generator=gen) * torch.tensor([0.3, 0.3, 0.3, 0.6, 0.6, 0.6])
    t[0] = 0
    t[1, 3:] *= 1e-5
    t[2, 3:] *= 1e-3
    t[3, 3:] *= 5.0  # large angle
    return t


def _se3_matrix_exp(t):
    w, z = t[:, 3:].double(), torch.zeros(t.shape[0], dtype=torch.float64)
    X = torch.zeros(t.shape[0], 4, 4, dtype=torch.float64)
    X[:, :3, :3] = torch.stack([z, -w[:, 2], w[:, 1], w[:, 2], z, -w[:, 0], -w[:, 1], w[:, 0], z], -1).reshape(-1, 3, 3)
    X[:, :3, 3] = t[:, :3].double()
    return torch.linalg.matrix_exp(X)[:, :3]


def test_oracle_se3_closed_form_is_the_matrix_exponential():
    t = _se3_tangents()
    torch.testing.assert_close(O.exp_map_se3(t).double(), _se3_matrix_exp(t), rtol=0, atol=2e-6)


def test_pixel_index_edge_cases():
    # truncation toward zero of the fp32 product; the largest fp32 below 1 stays inside for power-of-two sizes
    u = torch.tensor([[0.0, 0.0, 0.0], [0.999999, 0.5, 0.25], [np.nextafter(np.float32(1), np.float32(0)), 0.0, 0.0]])
    idx = O.pixel_indices(u, 128, 480, 640)
    assert idx.tolist() == [[0, 0, 0], [127, 240, 160], [127, 0, 0]]


# ---- GPU ------------------------------------------------------------------------------------------------------------------
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nv():
    import nerf_vo_b200 as nv

    assert torch.cuda.is_available()
    nv._lib.load()
    return nv


def _dataset(nv, g, use_normals=True, mode="off"):
    K, H, W, _ = g["frames_color"].shape
    dm = nv.DynamicDataManager(nv.DynamicDataManagerConfig(train_num_rays_per_batch=g["u"].shape[0], num_frames=K + 3, frame_height=H, frame_width=W,
                                                           use_normals=use_normals), device=DEV)
    ds = dm.train_dataset
    ds.camera_intrinsics[:K] = T(g["intrinsics"], DEV)
    ds.camera_extrinsics[:K] = T(g["extrinsics"], DEV)
    ds.frames_color[:K] = T(g["frames_color"], DEV)
    ds.frames_depth[:K] = T(g["frames_depth"], DEV)
    if use_normals:
        ds.frames_normal[:K] = T(g["frames_normal"], DEV)
    ds.num_active_frames = K
    if mode != "off":
        dm.camera_optimizer = nv.CameraOptimizerConfig(mode=mode).setup(num_cameras=K + 3, device=DEV)
        with torch.no_grad():
            dm.camera_optimizer.pose_adjustment[:K] = T(g["pose_adjustment"], DEV)
    return dm


@pytest.mark.gpu
def test_next_train_matches_reference_golden(nv, golden):
    g = golden("batch_prologue")
    dm = _dataset(nv, g)
    rb, batch = dm.next_train(0, u=T(g["u"], DEV))
    torch.cuda.synchronize()
    assert torch.equal(batch["indices"].cpu(), T(g["indices"])), "pixel indices must be bit-exact"
    assert torch.equal(rb.camera_indices.cpu(), T(g["camera_indices"]))
    assert torch.equal(batch["image"].cpu(), T(g["image"])) and torch.equal(batch["depth_image"].cpu(), T(g["depth_image"]))
    torch.testing.assert_close(batch["normal_image"].cpu(), T(g["normal_image"]), rtol=0, atol=1e-5)
    torch.testing.assert_close(rb.origins.cpu(), T(g["origins"]), rtol=0, atol=0)
    torch.testing.assert_close(rb.directions.cpu(), T(g["directions"]), rtol=0, atol=2e-6)
    torch.testing.assert_close(rb.metadata["directions_norm"].cpu(), T(g["directions_norm"]), rtol=2e-6, atol=0)
    torch.testing.assert_close(rb.pixel_area.cpu(), T(g["pixel_area"]), rtol=2e-4, atol=1e-9)  # a product of two differences of nearby unit vectors


@pytest.mark.gpu
def test_ray_generator_and_pixel_sampler_classes(nv, golden):
    g = golden("batch_prologue")
    dm = _dataset(nv, g)
    rb = dm.train_ray_generator(T(g["indices"], DEV))
    torch.testing.assert_close(rb.directions.cpu(), T(g["directions"]), rtol=0, atol=2e-6)
    torch.testing.assert_close(rb.origins.cpu(), T(g["origins"]), rtol=0, atol=0)
    # PixelSampler.sample over the reference-shaped dataset dict: same torch.rand stream => the reference's pixels
    torch.manual_seed(5)
    u = torch.rand((g["u"].shape[0], 3), device=DEV)
    torch.manual_seed(5)
    batch = dm.train_pixel_sampler.sample(dm.train_dataset.get_dataset())
    want = O.pixel_indices(u.cpu(), g["frames_color"].shape[0], g["frames_color"].shape[1], g["frames_color"].shape[2])
    assert torch.equal(batch["indices"].cpu(), want)
    c, y, x = want.unbind(-1)
    assert torch.equal(batch["image"].cpu(), T(g["frames_color"])[c, y, x])
    ref_n = O.next_train_batch(u.cpu(), T(g["intrinsics"]), T(g["extrinsics"]), T(g["frames_color"]), T(g["frames_depth"]), T(g["frames_normal"]))[1]["normal_image"]
    torch.testing.assert_close(batch["normal_image"].cpu(), ref_n, rtol=0, atol=1e-5)


@pytest.mark.gpu
def test_full_frame_rays(nv, golden):
    g = golden("batch_prologue")
    dm = _dataset(nv, g)
    K, H, W, _ = g["frames_color"].shape
    cam = 4
    rb = dm.train_dataset.cameras.generate_rays(cam)
    assert rb.origins.shape == (H, W, 3) and rb.pixel_area.shape == (H, W, 1)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    idx = torch.stack([torch.full_like(ys, cam), ys, xs], -1).reshape(-1, 3)
    want = O.generate_rays(idx, T(g["intrinsics"]), T(g["extrinsics"])[:, :3])
    torch.testing.assert_close(rb.directions.reshape(-1, 3).cpu(), want["directions"], rtol=0, atol=2e-6)
    torch.testing.assert_close(rb.origins.reshape(-1, 3).cpu(), want["origins"], rtol=0, atol=0)
    torch.testing.assert_close(rb.metadata["directions_norm"].reshape(-1, 1).cpu(), want["directions_norm"], rtol=2e-6, atol=0)
    assert torch.equal(rb.camera_indices.reshape(-1, 1).cpu(), want["camera_indices"])


@pytest.mark.gpu
def test_pose_correction_so3xr3_golden(nv, golden):
    g = golden("batch_prologue")
    dm = _dataset(nv, g, use_normals=False, mode="SO3xR3")
    K = g["frames_color"].shape[0]
    M = dm.camera_optimizer(torch.arange(K, device=DEV))
    torch.testing.assert_close(M.cpu(), T(g["pose_matrices_so3xr3"]), rtol=0, atol=1e-6)
    rb, _ = dm.next_train(0, u=T(g["u"], DEV))  # fused: correction applied inside the prologue kernel
    torch.testing.assert_close(rb.origins.cpu(), T(g["origins_so3xr3"]), rtol=0, atol=1e-6)
    torch.testing.assert_close(rb.directions.cpu(), T(g["directions_so3xr3"]), rtol=0, atol=2e-6)
    # unfused route (reference call order: RayGenerator, then CameraOptimizer.apply_to_raybundle) with autograd to the pose parameters
    rb2 = dm.train_ray_generator(T(g["indices"], DEV))
    dm.camera_optimizer.apply_to_raybundle(rb2)
    torch.testing.assert_close(rb2.directions.detach().cpu(), T(g["directions_so3xr3"]), rtol=0, atol=2e-6)
    ((rb2.origins * T(g["cot_origins"], DEV)).sum() + (rb2.directions * T(g["cot_directions"], DEV)).sum()).backward()
    got = dm.camera_optimizer.pose_adjustment.grad[:K].cpu()
    want = T(g["pose_adjustment_grad"])
    assert float((got - want).abs().max()) < 1e-4 * float(want.abs().max())
    # the fused prologue's backward entry point gives the same gradient from its saved raw directions
    out = nv.ops.batch_prologue(T(g["u"], DEV), K, dm.train_dataset.camera_intrinsics, dm.train_dataset.camera_extrinsics, dm.train_dataset.frames_color,
                                dm.train_dataset.frames_depth, None, dm.camera_optimizer.pose_adjustment, 1, want_raw_directions=True)
    dp = nv.ops.pose_correction_backward(out["camera_indices"], out["directions_raw"], T(g["cot_origins"], DEV), T(g["cot_directions"], DEV),
                                         dm.camera_optimizer.pose_adjustment.detach(), 1)
    assert float((dp[:K].cpu() - want).abs().max()) < 1e-4 * float(want.abs().max())


@pytest.mark.gpu
def test_pose_se3_vs_oracle(nv):
    t = _se3_tangents()
    tg = t.to(DEV).requires_grad_(True)
    M = nv.ops.pose_exp_map(tg, 2)
    torch.testing.assert_close(M.detach().cpu().double(), _se3_matrix_exp(t), rtol=0, atol=5e-6)
    gen = torch.Generator().manual_seed(8)
    G = torch.randn(t.shape[0], 3, 4, generator=gen)
    (M * G.to(DEV)).sum().backward()
    to = t.clone().requires_grad_(True)
    (O.exp_map_se3(to) * G).sum().backward()
    assert float((tg.grad.cpu() - to.grad).abs().max()) < 2e-4 * float(to.grad.abs().max())


@pytest.mark.gpu
def test_prologue_errors_and_edges(nv, golden):
    g = golden("batch_prologue")
    dm = _dataset(nv, g)
    ds = dm.train_dataset
    with pytest.raises(RuntimeError):
        nv.ops.batch_prologue(T(g["u"]), 7, ds.camera_intrinsics, ds.camera_extrinsics, ds.frames_color, ds.frames_depth)  # CPU tensor
    with pytest.raises(RuntimeError):
        nv.ops.batch_prologue(T(g["u"], DEV), 0, ds.camera_intrinsics, ds.camera_extrinsics, ds.frames_color, ds.frames_depth)  # no active frame
    with pytest.raises(RuntimeError):
        nv.ops.batch_prologue(T(g["u"], DEV), 7, ds.camera_intrinsics, ds.camera_extrinsics, ds.frames_color, ds.frames_depth, None, None, 1)  # mode without poses
    out = nv.ops.batch_prologue(torch.empty((0, 3), device=DEV), 7, ds.camera_intrinsics, ds.camera_extrinsics, ds.frames_color, ds.frames_depth)
    assert out["origins"].shape == (0, 3)
    # u one ulp below 1: stays in range (the reference would index out of bounds when the fp32 product rounds up)
    u = torch.full((5, 3), float(np.nextafter(np.float32(1), np.float32(0))), device=DEV)
    out = nv.ops.batch_prologue(u, 7, ds.camera_intrinsics, ds.camera_extrinsics, ds.frames_color, ds.frames_depth)
    K, H, W, _ = g["frames_color"].shape
    assert out["indices"].cpu().tolist() == [[K - 1, H - 1, W - 1]] * 5


@pytest.mark.gpu
def test_replica_sized_prologue_properties(nv):
    """BASELINE config 2 sizes (192 keyframes of 360x640 at NeRF-VO's training resolution, 4096 rays): size-independent properties —
    indices in range and equal to the oracle's, gathered pixels equal a torch gather, unit directions, norms > 1 (z = -1 in camera frame)."""
    K, H, W, B = 192, 360, 640, 4096
    gen = torch.Generator().manual_seed(0)
    dm = nv.DynamicDataManager(nv.DynamicDataManagerConfig(train_num_rays_per_batch=B, num_frames=K, frame_height=H, frame_width=W), device=DEV)
    ds = dm.train_dataset
    ds.camera_intrinsics[:] = torch.tensor([320.0, 320.0, 319.5, 179.5], device=DEV)
    q = torch.nn.functional.normalize(torch.randn(K, 4, generator=gen), dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1).reshape(K, 3, 3)
    ds.camera_extrinsics[:, :3, :3] = R.to(DEV)
    ds.camera_extrinsics[:, :3, 3] = (torch.rand(K, 3, generator=gen) - 0.5).to(DEV)
    ds.frames_color.uniform_()
    ds.frames_depth.uniform_()
    ds.frames_normal.normal_()
    ds.num_active_frames = K
    u = torch.rand(B, 3, generator=gen)
    rb, batch = dm.next_train(0, u=u.to(DEV))
    idx = batch["indices"]
    assert torch.equal(idx.cpu(), O.pixel_indices(u, K, H, W))
    c, yy, xx = idx.unbind(-1)
    assert torch.equal(batch["image"], ds.frames_color[c, yy, xx]) and torch.equal(batch["depth_image"], ds.frames_depth[c, yy, xx])
    torch.testing.assert_close(rb.directions.norm(dim=-1), torch.ones(B, device=DEV), rtol=0, atol=1e-6)
    assert bool((rb.metadata["directions_norm"] >= 1.0).all())
    want = O.generate_rays(idx.cpu(), ds.camera_intrinsics.cpu(), ds.camera_extrinsics[:, :3].cpu())
    torch.testing.assert_close(rb.directions.cpu(), want["directions"], rtol=0, atol=2e-6)
    n = torch.einsum("nij,nj->ni", R[c.cpu()].transpose(1, 2), ds.frames_normal[c, yy, xx].cpu())  # orthonormal R: R^-1 = R^T
    torch.testing.assert_close(batch["normal_image"].cpu(), (n + 1) / 2, rtol=0, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_trainer_fed_by_the_dataset(nv, graph):
    """MappingTrainer with a DynamicDataManager: prologue inside the step (and inside the CUDA graph); the loss is finite, goes down on a
    constant-colour scene, and the replayed graph draws fresh pixels every step."""
    from nerf_vo_b200.synthetic import synthetic_keyframes
    from nerf_vo_b200.trainer import MappingTrainer

    torch.manual_seed(0)
    K, H, W, B = 6, 48, 64, 512
    dm = nv.DynamicDataManager(nv.DynamicDataManagerConfig(train_num_rays_per_batch=B, num_frames=K, frame_height=H, frame_width=W), device=DEV)
    synthetic_keyframes(dm.train_dataset, seed=1, fx=40.0, fy=40.0)
    dm.train_dataset.frames_color[:] = torch.tensor([0.2, 0.5, 0.8], device=DEV)
    cfg = nv.NerfactoModelConfig(log2_hashmap_size=14)
    for a in cfg.proposal_net_args_list:
        a["log2_hashmap_size"] = 12
    model = nv.ExtendedNerfactoModel(cfg, num_train_data=K).to(DEV)
    tr = MappingTrainer(model, num_rays=B, use_cuda_graph=graph, datamanager=dm)
    tr.capture(warmup=2)
    losses, seen = [], []
    for _ in range(30):
        losses.append(float(tr.train_step()))
        seen.append(dm._u.clone())
    assert all(np.isfinite(losses))
    assert not torch.equal(seen[0], seen[1]), "every step must draw new pixels"
    rgb_first, rgb_last = np.mean(losses[:5]), np.mean(losses[-5:])
    assert rgb_last < rgb_first
