"""Evaluation frame loop.  CPU: the oracle's render_frame against the golden frame rendered by the reference
(oracle/make_golden.py:gen_frame_render).  GPU: nvo_b200.NerfstudioRenderer through the C ABI against the same golden and the oracle.
Tolerances: uint8 colour within 1 code value (a truncation boundary can flip under 1e-6 differences; fp16 production path: 1e-3
max-abs on rgb = 0.26 code values -> also within 1); depth: the median-depth sample index must agree on >= 95 % of the pixels
(test_model_eval_golden's criterion), elsewhere 1e-5; depth-scale sums 1e-6 relative; uint16 depth within 1 code value."""
import numpy as np
import pytest
import torch

import nerfacto_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def _cfg_and_params(g):
    cfg = O.ModelCfg(main_grid=O.GridCfg(log2_hashmap_size=int(g["main_log2"])),
                     prop_grids=(O.GridCfg(5, 16, 128, int(g["prop_log2"])), O.GridCfg(5, 16, 256, int(g["prop_log2"]))), num_images=int(g["K"]))
    P = {k[len("param."):]: T(v) for k, v in g.items() if k.startswith("param.")}
    return cfg, P


def test_oracle_render_frame_matches_reference_golden(golden):
    g, f = golden("model_step_small"), golden("frame_render_small")
    cfg, P = _cfg_and_params(g)
    H, W = f["depth"].shape
    rays = O.frame_rays(T(f["intrinsics"]), f["extrinsics"], H, W)
    torch.testing.assert_close(rays["directions"].view(H, W, 3), T(f["directions"]), rtol=0, atol=1e-7)
    torch.testing.assert_close(rays["directions_norm"].view(H, W, 1), T(f["directions_norm"]), rtol=1e-7, atol=0)
    color, depth = O.render_frame(P, cfg, T(f["intrinsics"]), f["extrinsics"], H, W, chunk=48)
    assert int(np.abs(color.astype(int) - f["color"].astype(int)).max()) <= 1 and float((color == f["color"]).mean()) > 0.99
    np.testing.assert_allclose(depth, f["depth"], rtol=0, atol=1e-6)
    sg, sp, cnt = O.depth_scale_sums(f["depth_gt"], f["depth"])
    assert cnt == int(f["mask_count"]) and abs((sg / cnt) / (sp / cnt) - float(f["scale"])) < 1e-6 * float(f["scale"])
    assert np.array_equal(O.depth_to_uint16(f["depth"], float(f["scale"]), 6553.5), f["depth16"])


# ---- GPU ------------------------------------------------------------------------------------------------------------------
DEV = "cuda:0"


@pytest.fixture(scope="module")
def nv():
    import nerf_vo_b200 as nv

    assert torch.cuda.is_available()
    nv._lib.load()
    return nv


def _model(nv, g, precision):
    cfg = nv.NerfactoModelConfig(log2_hashmap_size=int(g["main_log2"]), precision=precision)
    for a in cfg.proposal_net_args_list:
        a["log2_hashmap_size"] = int(g["prop_log2"])
    m = nv.ExtendedNerfactoModel(cfg, num_train_data=int(g["K"]))
    P = {k[len("param."):]: T(v) for k, v in g.items() if k.startswith("param.")}
    missing, unexpected = m.load_state_dict(P, strict=False)  # reference (torch layout) keys load directly
    assert not unexpected, unexpected
    return m.to(DEV).eval()


def _intr(f):
    fx, fy, cx, cy = [float(x) for x in f["intrinsics"]]
    H, W = f["depth"].shape
    return {"fx": fx, "fy": fy, "cx": cx, "cy": cy, "height": H, "width": W}


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_render_frame_matches_reference_golden(nv, golden, precision):
    g, f = golden("model_step_small"), golden("frame_render_small")
    r = nv.NerfstudioRenderer(model=_model(nv, g, precision), num_rays_per_chunk=48)
    ext = f["extrinsics"].copy()
    color, depth = r.render_frame(_intr(f), ext)
    assert np.array_equal(ext, f["extrinsics"]), "the caller's extrinsics must not be modified"
    assert color.dtype == np.uint8 and color.shape == f["color"].shape and depth.dtype == np.float32
    assert int(np.abs(color.astype(int) - f["color"].astype(int)).max()) <= 1
    bad = np.abs(depth - f["depth"]) > 1e-5
    assert float(bad.mean()) <= 0.05, float(bad.mean())
    # chunk-size independence: one 140-ray call and 48-ray chunks give the same bytes
    r2 = nv.NerfstudioRenderer(model=r.model, num_rays_per_chunk=1 << 16)
    color2, depth2 = r2.render_frame(_intr(f), f["extrinsics"].copy())
    assert np.array_equal(color, color2) and np.array_equal(depth, depth2)


@pytest.mark.gpu
def test_frame_finalize_and_depth_scale_vs_oracle(nv, golden):
    from nerf_vo_b200.nerf_renderer import depth_scale_sums, frame_finalize

    f = golden("frame_render_small")
    rgb, dep, dn = T(f["out.rgb"]).reshape(-1, 3), T(f["out.depth"]).reshape(-1, 1), T(f["directions_norm"]).reshape(-1, 1)
    scale = float(f["scale"])
    color, depth, d16 = frame_finalize(rgb.to(DEV), dep.to(DEV), dn.to(DEV), (scale, 6553.5))
    assert np.array_equal(color.cpu().numpy().reshape(f["color"].shape), f["color"]), "uint8 conversion must be bit-exact on equal inputs"
    assert np.array_equal(depth.cpu().numpy().reshape(f["depth"].shape), f["depth"])
    assert np.array_equal(d16.cpu().numpy().view(np.uint16).reshape(f["depth16"].shape), f["depth16"])
    s = depth_scale_sums(T(f["depth_gt"]).reshape(-1).to(DEV), depth).cpu().numpy()
    sg, sp, cnt = O.depth_scale_sums(f["depth_gt"], f["depth"])
    assert int(s[2]) == cnt and abs(s[0] - sg) < 1e-9 * sg and abs(s[1] - sp) < 1e-9 * sp
    # ragged sizes (n % 4 != 0) and the no-directions_norm branch
    gen = torch.Generator().manual_seed(0)
    for n in (1, 3, 5, 1023):
        rgb = torch.rand(n, 3, generator=gen)
        dep = torch.rand(n, 1, generator=gen) * 7
        c, d, h = frame_finalize(rgb.to(DEV), dep.to(DEV), None, (1.7, 1000.0))
        assert np.array_equal(c.cpu().numpy(), (rgb.numpy() * 255).astype(np.uint8))
        assert np.array_equal(d.cpu().numpy(), dep.numpy()[:, 0])
        assert np.array_equal(h.cpu().numpy().view(np.uint16), O.depth_to_uint16(dep.numpy()[:, 0], 1.7, 1000.0))


@pytest.mark.gpu
def test_replica_frame_properties(nv):
    """BASELINE config 5 size (1200x680 = 816 000 rays, forward only): accumulation in [0,1], depth >= 0 and finite, colour range, and the
    rows rendered as two separate shards equal the full render (row sharding = what N GPUs do)."""
    torch.manual_seed(0)
    m = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(), num_train_data=8).to(DEV).eval()
    with torch.no_grad():
        m.field.mlp_base.encoder.hash_table.normal_(0, 0.1)
    r = nv.NerfstudioRenderer(model=m)
    intr = {"fx": 600.0, "fy": 600.0, "cx": 599.5, "cy": 339.5, "height": 680, "width": 1200}
    ext = np.eye(4)
    full = r.render_frame_device(intr, ext, depth16_scales=(1.0, 6553.5))
    assert full["color"].shape == (680, 1200, 3) and full["depth"].shape == (680, 1200)
    assert bool(torch.isfinite(full["depth"]).all()) and float(full["depth"].min()) >= 0.0
    top, bot = r.render_frame_device(intr, ext, rows=(0, 340)), r.render_frame_device(intr, ext, rows=(340, 680))
    assert torch.equal(torch.cat([top["color"], bot["color"]]), full["color"])
    assert torch.equal(torch.cat([top["depth"], bot["depth"]]), full["depth"])
    d16 = full["depth16"].cpu().numpy().view(np.uint16)
    assert np.array_equal(d16, O.depth_to_uint16(full["depth"].cpu().numpy(), 1.0, 6553.5))
