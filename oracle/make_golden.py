"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE
(nerfstudio torch implementation under /root/reference) on seeded inputs.  Runs only in the build
container; the fixtures it writes are committed so the GPU box (no /root/reference) can check
both the oracle restatement and the CUDA path against the reference's own numbers.

    python oracle/make_golden.py            # rewrites tests/golden/

Fixtures use reduced hash-table sizes (2^12 main / 2^10 proposal) where table contents are stored, so
that files stay small; integer hash indices are additionally pinned at the full 2^19 / 2^17 sizes.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import nerfacto_oracle as O  # noqa: E402
import reference_harness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def edge_positions(n_rand: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n_rand, 3, generator=g)
    special = torch.tensor(
        [
            [0.0, 0.0, 0.0],
            [1.0, 1.0, 1.0],
            [0.5, 0.5, 0.5],
            [0.25, 0.75, 0.125],  # exact lattice points on several levels (ceil == floor)
            [1.0 / 16, 2.0 / 16, 3.0 / 16],
            [0.999999, 0.999999, 0.999999],
            [1e-7, 1e-7, 1e-7],
            [0.0, 0.5, 1.0],
            [0.3333333, 0.6666667, 0.1],
        ]
    )
    return torch.cat([special, x], 0)


def gen_hash_indices():
    rh.install()
    from nerfstudio.field_components.encodings import HashEncoding

    x = edge_positions(87, 11)
    out = {"x": x}
    for name, kw in (
        ("main", dict(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=19)),
        ("prop0", dict(num_levels=5, min_res=16, max_res=128, log2_hashmap_size=17)),
        ("prop1", dict(num_levels=5, min_res=16, max_res=256, log2_hashmap_size=17)),
        ("main21", dict(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=21)),
    ):
        enc = HashEncoding(implementation="torch", **kw)
        rec = []
        orig = enc.hash_fn

        def hooked(t, _orig=orig, _rec=rec):
            r = _orig(t)
            _rec.append(r.clone())
            return r

        enc.hash_fn = hooked
        with torch.no_grad():
            enc.pytorch_fwd(x)
        out[f"{name}_indices"] = torch.stack(rec, dim=-1)  # [N, L, 8] reference corner order
        out[f"{name}_scalings"] = enc.scalings.clone()
        del enc
    np.savez_compressed(os.path.join(OUT, "hash_indices.npz"), **_np(out))


def gen_hashgrid_small():
    rh.install()
    from nerfstudio.field_components.encodings import HashEncoding

    torch.manual_seed(3)
    out = {}
    for name, kw in (
        ("main", dict(num_levels=16, min_res=16, max_res=2048, log2_hashmap_size=12)),
        ("prop", dict(num_levels=5, min_res=16, max_res=256, log2_hashmap_size=10)),
    ):
        enc = HashEncoding(implementation="torch", **kw)
        with torch.no_grad():
            enc.hash_table.normal_(0, 0.1)
        x = edge_positions(247, 12).requires_grad_(True)
        y = enc(x)
        G = torch.randn(y.shape, generator=torch.Generator().manual_seed(4))
        (y * G).sum().backward()
        out.update({f"{name}_x": x, f"{name}_table": enc.hash_table, f"{name}_y": y, f"{name}_G": G, f"{name}_dtable": enc.hash_table.grad,
                    f"{name}_dx": x.grad, f"{name}_scalings": enc.scalings, f"{name}_log2_T": kw["log2_hashmap_size"]})
    np.savez_compressed(os.path.join(OUT, "hashgrid_small.npz"), **_np(out))


def gen_mlp():
    rh.install()
    from nerfstudio.field_components.mlp import MLP

    torch.manual_seed(5)
    out = {}
    cases = {
        "base": dict(in_dim=32, num_layers=2, layer_width=64, out_dim=16, out_activation=None),
        "head": dict(in_dim=63, num_layers=3, layer_width=64, out_dim=3, out_activation=torch.nn.Sigmoid()),
        "pred": dict(in_dim=27, num_layers=3, layer_width=64, out_dim=64, out_activation=None),
        "prop": dict(in_dim=10, num_layers=2, layer_width=16, out_dim=1, out_activation=None),
    }
    for name, kw in cases.items():
        m = MLP(implementation="torch", activation=torch.nn.ReLU(), **kw)
        x = torch.randn(200, kw["in_dim"]).requires_grad_(True)
        y = m(x)
        G = torch.randn(y.shape)
        (y * G).sum().backward()
        out[f"{name}_x"], out[f"{name}_y"], out[f"{name}_G"], out[f"{name}_dx"] = x, y, G, x.grad
        for i, layer in enumerate(m.layers):
            out[f"{name}_w{i}"], out[f"{name}_b{i}"] = layer.weight, layer.bias
            out[f"{name}_dw{i}"], out[f"{name}_db{i}"] = layer.weight.grad, layer.bias.grad
    np.savez_compressed(os.path.join(OUT, "mlp_cases.npz"), **_np(out))


def gen_model_step(B=64, K=8, main_log2=12, prop_log2=10):
    m = rh.build_reference_model(main_log2=main_log2, prop_log2=prop_log2, num_images=K, seed=0)
    torch.manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "hash_table" in n:
                p.normal_(0, 0.1)
    params = rh.reference_state(m)
    rays, targets = O.synthetic_rays(B, num_images=K, seed=5)
    # spread rays so that some samples land inside and some outside the unit cube
    field_rec = {}
    orig_field_forward = m.field.forward

    def field_forward(ray_samples, compute_normals=False):
        fo = orig_field_forward(ray_samples, compute_normals=compute_normals)
        for k, v in fo.items():
            field_rec[str(k.value if hasattr(k, "value") else k)] = v.detach().clone()
        return fo

    m.field.forward = field_forward
    step = 300  # anneal exponent != 1 (NS/models/nerfacto.py:260-270)
    out, ld, jit, ss_calls = rh.reference_step(m, rays, targets, step=step)
    G = {"B": B, "K": K, "main_log2": main_log2, "prop_log2": prop_log2, "anneal": O.anneal_value(step)}
    for k, v in rays.items():
        G[f"rays.{k}"] = v
    for k, v in targets.items():
        G[f"targets.{k}"] = v
    for k, v in params.items():
        G[f"param.{k}"] = v
    for n, p in m.named_parameters():
        if n in params:
            G[f"grad.{n}"] = p.grad if p.grad is not None else torch.zeros_like(p)
    for i, j in enumerate(jit):
        G[f"jitter.{i}"] = j
    for k in ("rgb", "accumulation", "depth", "expected_depth", "normals", "pred_normals", "prop_depth_0", "prop_depth_1"):
        G[f"out.{k}"] = out[k]
    for i, (w, rs) in enumerate(zip(out["weights_list"], out["ray_samples_list"])):
        G[f"level{i}.weights"] = w[..., 0]
        G[f"level{i}.starts"] = rs.frustums.starts[..., 0]
        G[f"level{i}.ends"] = rs.frustums.ends[..., 0]
        G[f"level{i}.sdist"] = torch.cat([rs.spacing_starts[..., 0], rs.spacing_ends[..., -1:, 0]], -1)
    for k, v in field_rec.items():
        G[f"field.{k}"] = v
    for k, v in ld.items():
        G[f"loss.{k}"] = v
    # integer intermediates, in call order: pdf inds (2), median idx final, median idx prop0, prop1, interlevel lo/hi x2
    names = ["pdf_inds_1", "pdf_inds_2", "median_idx_final", "median_idx_prop0", "median_idx_prop1",
             "inter_lo_0", "inter_hi_0", "inter_lo_1", "inter_hi_1"]
    assert len(ss_calls) == len(names), len(ss_calls)
    for n, (_, _, r) in zip(names, ss_calls):
        G[f"int.{n}"] = r
    np.savez_compressed(os.path.join(OUT, "model_step_small.npz"), **_np(G))

    # eval-mode render of the same parameters (config 5 path: NS/models/base_model.py:164-192)
    from nerfstudio.cameras.rays import RayBundle

    m.eval()
    m.field.forward = orig_field_forward
    rb = RayBundle(origins=rays["origins"].clone(), directions=rays["directions"].clone(), pixel_area=rays["pixel_area"].clone(),
                   camera_indices=rays["camera_indices"].clone(), metadata={"directions_norm": rays["directions_norm"].clone()})
    with torch.no_grad():
        eo = m(rb)
    E = {f"out.{k}": eo[k] for k in ("rgb", "accumulation", "depth", "expected_depth", "normals", "pred_normals", "prop_depth_0", "prop_depth_1")}
    np.savez_compressed(os.path.join(OUT, "model_eval_small.npz"), **_np(E))


def gen_ray_ops():
    """Standalone sampler / weights / renderer / loss known-answers from the reference classes."""
    rh.install()
    from nerfstudio.cameras.rays import Frustums, RayBundle, RaySamples
    from nerfstudio.model_components import losses as RL
    from nerfstudio.model_components.ray_samplers import PDFSampler, UniformLinDispPiecewiseSampler
    from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, NormalsRenderer, RGBRenderer

    torch.manual_seed(21)
    B = 33
    rays, _ = O.synthetic_rays(B, num_images=4, seed=8)
    G = {}
    for mode in ("train", "eval"):
        rb = RayBundle(origins=rays["origins"], directions=rays["directions"], pixel_area=rays["pixel_area"], camera_indices=rays["camera_indices"],
                       nears=torch.ones(B, 1) * (0.05 if mode == "train" else 0.0), fars=torch.ones(B, 1) * 1000.0)
        us = UniformLinDispPiecewiseSampler(single_jitter=True)
        ps = PDFSampler(include_original=False, single_jitter=True)
        (us.train(), ps.train()) if mode == "train" else (us.eval(), ps.eval())
        with rh.record_rand() as rr, rh.record_searchsorted() as rs:
            s0 = us(rb, num_samples=256)
            dens = torch.rand(B, 256, 1) * 40 * (torch.rand(B, 256, 1) > 0.7)
            dens[0] = 0.0  # a ray with zero density everywhere
            dens[1] = 1e4  # saturated ray
            w0 = s0.get_weights(dens)
            s1 = ps(rb, s0, w0, num_samples=96)
        G[f"{mode}.nears"] = rb.nears
        G[f"{mode}.s0_starts"], G[f"{mode}.s0_ends"] = s0.frustums.starts[..., 0], s0.frustums.ends[..., 0]
        G[f"{mode}.s0_sdist"] = torch.cat([s0.spacing_starts[..., 0], s0.spacing_ends[..., -1:, 0]], -1)
        G[f"{mode}.density0"], G[f"{mode}.w0"] = dens[..., 0], w0[..., 0]
        G[f"{mode}.s1_starts"], G[f"{mode}.s1_ends"] = s1.frustums.starts[..., 0], s1.frustums.ends[..., 0]
        G[f"{mode}.s1_sdist"] = torch.cat([s1.spacing_starts[..., 0], s1.spacing_ends[..., -1:, 0]], -1)
        G[f"{mode}.pdf_inds"] = rs.calls[0][2]
        G[f"{mode}.pdf_cdf"], G[f"{mode}.pdf_u"] = rs.calls[0][0][0], rs.calls[0][0][1]
        if mode == "train":
            jv = [v for v in rr.values if tuple(v.shape) == (B, 1)]
            G["train.jitter0"], G["train.jitter1"] = jv[0], jv[1]
        # renderers on the 96-sample level
        rgb = torch.rand(B, 96, 3)
        nrm = torch.nn.functional.normalize(torch.randn(B, 96, 3), dim=-1)
        d1 = torch.rand(B, 96, 1) * 30 * (torch.rand(B, 96, 1) > 0.5)
        d1[0] = 0.0
        w1 = s1.get_weights(d1)
        r_rgb = RGBRenderer(background_color="last_sample")
        (r_rgb.train() if mode == "train" else r_rgb.eval())
        with rh.record_searchsorted() as rs2:
            med = DepthRenderer(method="median")(weights=w1, ray_samples=s1)
        G[f"{mode}.density1"], G[f"{mode}.w1"], G[f"{mode}.rgb_samples"], G[f"{mode}.normal_samples"] = d1[..., 0], w1[..., 0], rgb, nrm
        G[f"{mode}.rgb"] = r_rgb(rgb=rgb, weights=w1)
        G[f"{mode}.median_depth"], G[f"{mode}.median_idx"] = med, rs2.calls[0][2]
        G[f"{mode}.expected_depth"] = DepthRenderer(method="expected")(weights=w1, ray_samples=s1)
        G[f"{mode}.accumulation"] = AccumulationRenderer()(weights=w1)
        G[f"{mode}.normals"] = NormalsRenderer()(normals=nrm, weights=w1)
        if mode == "train":
            G["train.interlevel"] = RL.interlevel_loss([w0, w1], [s0, s1])
            G["train.distortion"] = RL.distortion_loss([w0, w1], [s0, s1])
            dgt = torch.rand(B, 1) * 3 * (torch.rand(B, 1) > 0.2)
            G["train.depth_gt"] = dgt
            G["train.depth_loss"] = RL.depth_loss(weights=w1, ray_samples=s1, termination_depth=dgt, predicted_depth=med, sigma=torch.tensor([0.001]),
                                                  directions_norm=rays["directions_norm"], is_euclidean=False, depth_loss_type=RL.DepthLossType.DS_NERF)
            ngt = torch.randn(B, 3)
            G["train.normal_gt"] = ngt
            G["train.normal_loss"] = RL.monosdf_normal_loss((G["train.normals"] + 1) / 2, ngt)
    for k, v in rays.items():
        G[f"rays.{k}"] = v
    np.savez_compressed(os.path.join(OUT, "ray_ops.npz"), **_np(G))


def gen_field_enc():
    """SH / posenc / contraction / trunc_exp known answers."""
    rh.install()
    from nerfstudio.field_components.activations import trunc_exp
    from nerfstudio.field_components.encodings import NeRFEncoding, SHEncoding
    from nerfstudio.field_components.spatial_distortions import SceneContraction

    torch.manual_seed(31)
    d = torch.nn.functional.normalize(torch.randn(50, 3), dim=-1)
    p = torch.cat([torch.randn(40, 3) * 0.5, torch.randn(40, 3) * 30, torch.tensor([[1.0, 0, 0], [0, -1.0, 0.5], [0, 0, 0], [0.999, 0.2, -0.1]])])
    x = torch.randn(64).mul(8).requires_grad_(True)
    y = trunc_exp(x)
    y.backward(torch.ones_like(y))
    G = {"dirs": d, "sh": SHEncoding(levels=4, implementation="torch")((d + 1) / 2), "pos": p,
         "posenc": NeRFEncoding(in_dim=3, num_frequencies=2, min_freq_exp=0, max_freq_exp=1, implementation="torch")(p),
         "contracted": SceneContraction(order=float("inf"))(p), "te_x": x, "te_y": y, "te_dx": x.grad}
    np.savez_compressed(os.path.join(OUT, "field_enc.npz"), **_np(G))


def gen_batch_prologue(B=193, K=7, H=24, W=40):
    """Pixel sampling + pixel gather + ray generation + camera-pose correction, through the reference's own
    PixelSampler.sample (NS/data/pixel_samplers.py:170-219,300-317), RayGenerator.forward
    (NS/model_components/ray_generators.py:40-57 -> Cameras.generate_rays, NS/cameras/cameras.py:503-912) and
    CameraOptimizer.apply_to_raybundle (NS/cameras/camera_optimizers.py:108-147, mode SO3xR3).  The dataset dict is
    built exactly as DynamicDataset.get_dataset does (nerf_vo/mapping/nerfstudio_utils.py:133-155: the per-frame
    `linalg.solve(R, n)` normal transform over the whole frames); that module itself does not import here (its
    datamanager base pulls in the dataparser registry), so those 12 lines are executed verbatim on the same tensors."""
    rh.install()
    from nerfstudio.cameras import camera_utils
    from nerfstudio.cameras.camera_optimizers import CameraOptimizerConfig
    from nerfstudio.cameras.cameras import Cameras, CameraType
    from nerfstudio.data.pixel_samplers import PixelSamplerConfig
    from nerfstudio.model_components.ray_generators import RayGenerator

    g = torch.Generator().manual_seed(77)
    intr = torch.stack([torch.rand(K, generator=g) * 10 + 20, torch.rand(K, generator=g) * 10 + 22, W / 2 - 0.5 + torch.rand(K, generator=g),
                        H / 2 - 0.5 + torch.rand(K, generator=g)], dim=-1)
    q = torch.nn.functional.normalize(torch.randn(K, 4, generator=g), dim=-1)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=-1).reshape(K, 3, 3)
    ext = torch.eye(4).repeat(K, 1, 1)
    ext[:, :3, :3] = R * (0.5 + torch.rand(K, 1, 1, generator=g))  # scaled rotations: the scene normalisation scales poses
    ext[:, :3, 3] = torch.rand(K, 3, generator=g) - 0.5
    color = torch.rand(K, H, W, 3, generator=g)
    depth = torch.rand(K, H, W, 1, generator=g) * 4
    normal = torch.nn.functional.normalize(torch.randn(K, H, W, 3, generator=g), dim=-1)
    # DynamicDataset.get_dataset (nerfstudio_utils.py:133-155)
    frames_normal = (torch.linalg.solve(ext[:, :3, :3], normal.permute(0, 3, 1, 2).reshape(K, 3, H * W)).reshape(K, 3, H, W).permute(0, 2, 3, 1) + 1) / 2
    dataset = {"image_idx": torch.arange(0, K, dtype=torch.long), "image": color, "depth_image": depth, "normal_image": frames_normal}
    cameras = Cameras(fx=intr[:, 0], fy=intr[:, 1], cx=intr[:, 2], cy=intr[:, 3],
                      distortion_params=camera_utils.get_distortion_params(k1=0, k2=0, k3=0, k4=0, p1=0, p2=0), height=H, width=W,
                      camera_to_worlds=ext[:, :3], camera_type=CameraType.PERSPECTIVE)
    sampler = PixelSamplerConfig().setup(num_rays_per_batch=B)
    torch.manual_seed(5)
    with rh.record_rand() as rec:
        batch = sampler.sample(dataset)
    assert len(rec.values) == 1 and rec.values[0].shape == (B, 3)
    rb = RayGenerator(cameras)(batch["indices"])
    G = {"intrinsics": intr, "extrinsics": ext, "frames_color": color, "frames_depth": depth, "frames_normal": normal, "u": rec.values[0],
         "indices": batch["indices"], "image": batch["image"], "depth_image": batch["depth_image"], "normal_image": batch["normal_image"],
         "origins": rb.origins, "directions": rb.directions, "pixel_area": rb.pixel_area, "camera_indices": rb.camera_indices,
         "directions_norm": rb.metadata["directions_norm"]}
    opt = CameraOptimizerConfig(mode="SO3xR3").setup(num_cameras=K, device="cpu")
    adj = torch.randn(K, 6, generator=g) * torch.tensor([0.05, 0.05, 0.05, 0.2, 0.2, 0.2])
    adj[0] = 0.0  # zero tangent: the clamp(nrms, 1e-4) branch
    adj[1, 3:] = torch.tensor([1e-3, -2e-3, 5e-4])  # below the clamp
    with torch.no_grad():
        opt.pose_adjustment.copy_(adj)
    G["pose_adjustment"] = adj
    G["pose_matrices_so3xr3"] = opt(torch.arange(K)).detach()
    o0, d0 = rb.origins.clone(), rb.directions.clone()
    opt.apply_to_raybundle(rb)
    G["origins_so3xr3"], G["directions_so3xr3"] = rb.origins.detach(), rb.directions.detach()
    # gradient of a fixed linear functional of the corrected rays w.r.t. the pose parameters (autograd through the reference code)
    co, cd = torch.randn(B, 3, generator=g), torch.randn(B, 3, generator=g)
    ((rb.origins * co).sum() + (rb.directions * cd).sum()).backward()
    G["cot_origins"], G["cot_directions"], G["pose_adjustment_grad"] = co, cd, opt.pose_adjustment.grad.clone()
    np.savez_compressed(os.path.join(OUT, "batch_prologue.npz"), **_np(G))


def gen_frame_render(H=10, W=14, K=8, main_log2=12, prop_log2=10):
    """Evaluation frame render through the reference: the body of NerfstudioRenderer.render_frame
    (evaluation/nerf_renderer.py:132-168 — that module imports open3d / the instant-ngp build, so its lines are executed here on the
    reference's own Cameras.generate_rays(camera_indices=0, keep_shape=True) and Model.get_outputs_for_camera_ray_bundle,
    NS/models/base_model.py:164-192), plus the depth-scale alignment arithmetic of evaluation/renderer.py:79-97 and the uint16
    depth conversion of :113-121.  Model parameters are those of model_step_small.npz (same seeds), so they are not stored again."""
    rh.install()
    from nerfstudio.cameras import camera_utils
    from nerfstudio.cameras.cameras import Cameras, CameraType

    m = rh.build_reference_model(main_log2=main_log2, prop_log2=prop_log2, num_images=K, seed=0)
    torch.manual_seed(1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "hash_table" in n:
                p.normal_(0, 0.1)
    m.eval()
    m.config.eval_num_rays_per_chunk = 48  # ragged chunks: 140 rays = 48 + 48 + 44
    intr = {"fx": 9.0, "fy": 9.5, "cx": 6.4, "cy": 5.2, "height": H, "width": W}
    g = torch.Generator().manual_seed(21)
    q = torch.nn.functional.normalize(torch.randn(4, generator=g), dim=-1)
    w, x, y, z = q.tolist()
    ext = np.eye(4)
    ext[:3, :3] = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    ext[:3, 3] = [0.1, -0.2, 0.05]
    camera_extrinsics = ext.copy()
    # ---- evaluation/nerf_renderer.py:134-168 ----
    camera_extrinsics[0:3, 1:3] *= -1
    cameras = Cameras(fx=intr["fx"], fy=intr["fy"], cx=intr["cx"], cy=intr["cy"],
                      distortion_params=camera_utils.get_distortion_params(k1=0, k2=0, k3=0, k4=0, p1=0, p2=0), height=intr["height"], width=intr["width"],
                      camera_to_worlds=torch.Tensor(camera_extrinsics).unsqueeze(0)[:, :3], camera_type=CameraType.PERSPECTIVE)
    camera_ray_bundle = cameras.generate_rays(camera_indices=0, keep_shape=True)
    with torch.no_grad():
        outputs = m.get_outputs_for_camera_ray_bundle(camera_ray_bundle)
    color = (outputs["rgb"].cpu().numpy() * 255).astype(np.uint8)
    depth = (outputs["depth"] / camera_ray_bundle.metadata["directions_norm"]).cpu().numpy()[..., 0]
    # ---- evaluation/renderer.py:79-97 (one keyframe) and :113-121 ----
    frame_depth_gt = (torch.rand(H, W, generator=g) * 6).numpy()
    frame_depth_gt[0, :3] = 0
    mask = (frame_depth_gt > 0) * (depth > 0) * (frame_depth_gt < 5) * (depth < 5)
    scale = frame_depth_gt[mask].mean() / depth[mask].mean()
    depth16 = (depth.copy() * np.float32(scale) * np.float32(6553.5)).astype(np.uint16)
    G = {"intrinsics": np.array([intr["fx"], intr["fy"], intr["cx"], intr["cy"]], dtype=np.float32), "extrinsics": ext, "color": color, "depth": depth,
         "out.rgb": outputs["rgb"], "out.depth": outputs["depth"], "out.accumulation": outputs["accumulation"],
         "directions_norm": camera_ray_bundle.metadata["directions_norm"], "origins": camera_ray_bundle.origins, "directions": camera_ray_bundle.directions,
         "depth_gt": frame_depth_gt, "mask_count": int(mask.sum()), "scale": np.float64(scale), "depth16": depth16}
    np.savez_compressed(os.path.join(OUT, "frame_render_small.npz"), **_np(G))


def gen_checkpoint(K=8, main_log2=12, prop_log2=10, B=64):
    """A checkpoint dict exactly as NS/engine/trainer.py:436-447 writes it, from the UNMODIFIED reference model after two real optimizer
    steps (one torch.optim.Adam per parameter group, NS/engine/optimizers.py:138-150): keys / shapes / dtypes as JSON, tensors as .pt."""
    import json

    model = rh.build_reference_model(main_log2=main_log2, prop_log2=prop_log2, num_images=K, seed=0)
    groups = model.get_param_groups()
    groups = {k: v for k, v in groups.items() if len(v) > 0}
    opts = {k: torch.optim.Adam(v, lr=1e-2, eps=1e-15) for k, v in groups.items()}
    rays, targets = O.synthetic_rays(B, num_images=K, seed=11)
    for it in range(2):
        for o in opts.values():
            o.zero_grad(set_to_none=True)
        out, ld, jit, _ = rh.reference_step(model, rays, targets, step=it)
        for o in opts.values():
            o.step()
    ckpt = {"step": 1, "pipeline": {"_model." + k: v.detach().clone() for k, v in model.state_dict().items()},
            "optimizers": {k: o.state_dict() for k, o in opts.items()}, "schedulers": {}, "scalers": {}}
    torch.save(ckpt, os.path.join(OUT, "reference_checkpoint_small.pt"))
    meta = {"keys": [[k, list(v.shape), str(v.dtype)] for k, v in ckpt["pipeline"].items()],
            "groups": {k: [len(o.state_dict()["param_groups"][0]["params"]), sorted(int(i) for i in o.state_dict()["state"].keys())] for k, o in opts.items()}}
    with open(os.path.join(OUT, "reference_checkpoint_keys.json"), "w") as f:
        json.dump(meta, f, indent=0)
    print("checkpoint:", len(meta["keys"]), "keys,", {k: v[0] for k, v in meta["groups"].items()})


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_hash_indices()
    gen_hashgrid_small()
    gen_mlp()
    gen_field_enc()
    gen_ray_ops()
    gen_model_step()
    gen_batch_prologue()
    gen_frame_render()
    gen_checkpoint()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
