"""The CPU oracle restatement (oracle/nerfacto_oracle.py) against golden vectors produced by the
UNMODIFIED reference (oracle/make_golden.py).  Runs everywhere (no GPU, no /root/reference)."""
import numpy as np
import torch

import nerfacto_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_hash_indices_bit_exact(golden):
    g = golden("hash_indices")
    x = T(g["x"])
    for name, gc in (("main", O.GridCfg()), ("prop0", O.GridCfg(5, 16, 128, 17)), ("prop1", O.GridCfg(5, 16, 256, 17)),
                     ("main21", O.GridCfg(log2_hashmap_size=21))):
        sc = O.level_scalings(gc)
        assert torch.equal(sc, T(g[f"{name}_scalings"]))
        idx = O.hash_indices(x, sc, gc.log2_hashmap_size)
        assert torch.equal(idx, T(g[f"{name}_indices"]))


def test_main_scalings_known_answer():
    # fp32 evaluation of floor(16 * g**l) (note 2047, not 2048)
    assert O.level_scalings(O.GridCfg()).tolist() == [16, 22, 30, 42, 58, 80, 111, 153, 212, 294, 406, 561, 776, 1072, 1482, 2047]
    assert O.level_scalings(O.GridCfg(5, 16, 128, 17)).tolist() == [16, 26, 45, 76, 128]
    assert O.level_scalings(O.GridCfg(5, 16, 256, 17)).tolist() == [16, 32, 64, 128, 256]


def test_hashgrid_small_fwd_bwd(golden):
    g = golden("hashgrid_small")
    for name, gc in (("main", O.GridCfg(log2_hashmap_size=12)), ("prop", O.GridCfg(5, 16, 256, 10))):
        x = T(g[f"{name}_x"]).requires_grad_(True)
        tab = T(g[f"{name}_table"]).requires_grad_(True)
        y = O.hash_encode(x, tab, O.level_scalings(gc), gc.log2_hashmap_size)
        assert torch.equal(y.detach(), T(g[f"{name}_y"]))
        (y * T(g[f"{name}_G"])).sum().backward()
        torch.testing.assert_close(tab.grad, T(g[f"{name}_dtable"]), rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(x.grad, T(g[f"{name}_dx"]), rtol=1e-5, atol=1e-5)


def test_mlp_cases(golden):
    g = golden("mlp_cases")
    for name, n, act in (("base", 2, None), ("head", 3, "sigmoid"), ("pred", 3, None), ("prop", 2, None)):
        ws = [T(g[f"{name}_w{i}"]).requires_grad_(True) for i in range(n)]
        bs = [T(g[f"{name}_b{i}"]).requires_grad_(True) for i in range(n)]
        x = T(g[f"{name}_x"]).requires_grad_(True)
        y = O.mlp_forward(x, ws, bs, act)
        torch.testing.assert_close(y.detach(), T(g[f"{name}_y"]), rtol=1e-6, atol=1e-6)
        (y * T(g[f"{name}_G"])).sum().backward()
        torch.testing.assert_close(x.grad, T(g[f"{name}_dx"]), rtol=1e-5, atol=1e-6)
        for i in range(n):
            torch.testing.assert_close(ws[i].grad, T(g[f"{name}_dw{i}"]), rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(bs[i].grad, T(g[f"{name}_db{i}"]), rtol=1e-5, atol=1e-5)


def test_field_encodings(golden):
    g = golden("field_enc")
    assert torch.equal(O.sh_deg4((T(g["dirs"]) + 1) / 2), T(g["sh"]))
    assert torch.equal(O.posenc_2freq(T(g["pos"])), T(g["posenc"]))
    assert torch.equal(O.contract_linf(T(g["pos"])), T(g["contracted"]))
    x = T(g["te_x"]).requires_grad_(True)
    y = O.trunc_exp(x)
    y.backward(torch.ones_like(y))
    assert torch.equal(y.detach(), T(g["te_y"])) and torch.equal(x.grad, T(g["te_dx"]))


def test_ray_ops(golden):
    g = golden("ray_ops")
    B = g["rays.origins"].shape[0]
    fars = torch.ones(B, 1) * 1000.0
    for mode in ("train", "eval"):
        nears = T(g[f"{mode}.nears"])
        j0 = T(g["train.jitter0"]) if mode == "train" else None
        j1 = T(g["train.jitter1"]) if mode == "train" else None
        sb = O.uniform_spacing_bins(B, 256, j0)
        if sb.shape[0] != B:
            sb = sb.expand(B, -1)
        assert torch.equal(sb, T(g[f"{mode}.s0_sdist"]))
        eb = O.spacing_to_euclidean(sb, nears, fars)
        assert torch.equal(eb[:, :-1], T(g[f"{mode}.s0_starts"])) and torch.equal(eb[:, 1:], T(g[f"{mode}.s0_ends"]))
        w0 = O.get_weights(eb[:, 1:] - eb[:, :-1], T(g[f"{mode}.density0"]))
        assert torch.equal(w0, T(g[f"{mode}.w0"]))
        res = O.pdf_resample(w0, sb, 96, j1)
        assert torch.equal(res["inds"], T(g[f"{mode}.pdf_inds"]))
        assert torch.equal(res["cdf"], T(g[f"{mode}.pdf_cdf"])) and torch.equal(res["u"], T(g[f"{mode}.pdf_u"]))
        assert torch.equal(res["bins"], T(g[f"{mode}.s1_sdist"]))
        e1 = O.spacing_to_euclidean(res["bins"], nears, fars)
        s1, e1 = e1[:, :-1], e1[:, 1:]
        assert torch.equal(s1, T(g[f"{mode}.s1_starts"]))
        w1 = O.get_weights(e1 - s1, T(g[f"{mode}.density1"]))
        assert torch.equal(w1, T(g[f"{mode}.w1"]))
        rgb = O.render_rgb(T(g[f"{mode}.rgb_samples"]), w1, training=(mode == "train"))
        assert torch.equal(rgb, T(g[f"{mode}.rgb"]))
        md, mi = O.render_depth_median(w1, s1, e1)
        # golden holds the raw searchsorted result; the reference clamps it afterwards (renderers.py:359)
        assert torch.equal(mi, T(g[f"{mode}.median_idx"]).clamp(0, 95)) and torch.equal(md, T(g[f"{mode}.median_depth"]))
        assert torch.equal(O.render_depth_expected(w1, s1, e1), T(g[f"{mode}.expected_depth"]))
        assert torch.equal(O.render_accumulation(w1), T(g[f"{mode}.accumulation"]))
        nr = O.render_normals(T(g[f"{mode}.normal_samples"]), w1)
        torch.testing.assert_close(nr * 2 - 1, T(g[f"{mode}.normals"]), rtol=0, atol=2e-7)
        if mode == "train":
            torch.testing.assert_close(O.interlevel_loss([w0, w1], [sb, res["bins"]]), T(g["train.interlevel"]), rtol=1e-6, atol=0)
            torch.testing.assert_close(O.distortion_loss(w1, res["bins"]), T(g["train.distortion"]), rtol=1e-6, atol=0)
            dl = O.ds_nerf_depth_loss(w1, s1, e1, T(g["train.depth_gt"]), T(g["rays.directions_norm"]), 0.001)
            torch.testing.assert_close(dl, T(g["train.depth_loss"]), rtol=1e-6, atol=0)
            torch.testing.assert_close(O.monosdf_normal_loss(T(g["train.normals"] + 1) / 2, T(g["train.normal_gt"])), T(g["train.normal_loss"]), rtol=1e-6, atol=0)


def small_cfg(g):
    return O.ModelCfg(main_grid=O.GridCfg(log2_hashmap_size=int(g["main_log2"])),
                      prop_grids=(O.GridCfg(5, 16, 128, int(g["prop_log2"])), O.GridCfg(5, 16, 256, int(g["prop_log2"]))), num_images=int(g["K"]))


def load_step(g):
    P = {k[len("param."):]: T(v).clone() for k, v in g.items() if k.startswith("param.")}
    rays = {k[len("rays."):]: T(v) for k, v in g.items() if k.startswith("rays.")}
    targets = {k[len("targets."):]: T(v) for k, v in g.items() if k.startswith("targets.")}
    jit = [T(g[f"jitter.{i}"]) for i in range(3)]
    return P, rays, targets, jit


def test_model_step_small(golden):
    g = golden("model_step_small")
    cfg = small_cfg(g)
    P, rays, targets, jit = load_step(g)
    P = {k: v.requires_grad_(True) for k, v in P.items()}
    out, L, total = O.mapping_step(P, cfg, rays, targets, jit, anneal=float(g["anneal"]))
    for k in ("rgb", "accumulation", "depth", "expected_depth", "normals", "pred_normals", "prop_depth_0", "prop_depth_1"):
        torch.testing.assert_close(out[k], T(g[f"out.{k}"]), rtol=0, atol=1e-6, msg=k)
    for i in range(3):
        assert torch.equal(out["weights_list"][i].detach(), T(g[f"level{i}.weights"]))
        assert torch.equal(out["sdist_list"][i], T(g[f"level{i}.sdist"]))
        assert torch.equal(out["starts_list"][i], T(g[f"level{i}.starts"]))
    assert torch.equal(out["pdf_aux"][0]["inds"], T(g["int.pdf_inds_1"])) and torch.equal(out["pdf_aux"][1]["inds"], T(g["int.pdf_inds_2"]))
    assert torch.equal(out["depth_index"], T(g["int.median_idx_final"]).clamp(0, 47))
    torch.testing.assert_close(out["field"]["density"].detach(), T(g["field.density"])[..., 0], rtol=1e-6, atol=0)
    torch.testing.assert_close(out["field"]["rgb"].detach(), T(g["field.rgb"]), rtol=0, atol=1e-6)
    torch.testing.assert_close(out["field"]["normals"], T(g["field.normals"]), rtol=0, atol=1e-5)
    for k, v in L.items():
        torch.testing.assert_close(v.detach(), T(g[f"loss.{k}"]).reshape(v.shape), rtol=1e-5, atol=1e-12, msg=k)
    for k, p in P.items():
        ref = T(g[f"grad.{k}"])
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-9, msg=k)


def test_model_eval_small(golden):
    g = golden("model_step_small")
    e = golden("model_eval_small")
    cfg = small_cfg(g)
    P, rays, _, _ = load_step(g)
    with torch.no_grad():
        out = O.mapping_forward(P, cfg, rays["origins"], rays["directions"], rays["camera_indices"], None, float(g["anneal"]), training=False)  # sampler keeps its last anneal
    for k in ("rgb", "accumulation", "depth", "expected_depth", "normals", "pred_normals", "prop_depth_0", "prop_depth_1"):
        torch.testing.assert_close(out[k], T(e[f"out.{k}"]), rtol=0, atol=1e-6, msg=k)
