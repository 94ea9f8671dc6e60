"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there): configs[3]'s 2^21-row main table and configs[2]'s
65 536-ray batch.  What must hold whatever the size: trilinear weights sum to one (the scatter conserves every level's gradient mass), the encoding
is linear in the table, hash rows stay inside their level, sample bins are sorted, weights are a sub-probability distribution per ray and the
accumulation is their sum."""
import numpy as np
import pytest
import torch

import nerfacto_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rows_to_tmf(a):
    n, k = a.shape
    tiles = (n + 127) // 128
    pad = torch.zeros(tiles * 128, k, dtype=a.dtype, device=a.device)
    pad[:n] = a
    return pad.view(tiles, 128, k).permute(0, 2, 1).contiguous().view(-1)


def test_config4_table_2pow21_linearity_and_scatter_conservation():
    import nerf_vo_b200 as nv

    L, log2, n = 16, 21, 262144
    sc = O.level_scalings(O.GridCfg(log2_hashmap_size=log2))
    spec = nv.ops.GridSpec(L, log2, tuple(float(s) for s in sc))
    g = torch.Generator(device=DEV).manual_seed(5)
    o = torch.rand(n // 16, 3, device=DEV, generator=g)
    d = torch.randn(n // 16, 3, device=DEV, generator=g) * 0.01
    x = (o[:, None] + d[:, None] * torch.linspace(0, 1, 16, device=DEV)[None, :, None]).reshape(-1, 3).clamp(1e-4, 1 - 1e-4).contiguous()
    # hash rows stay inside their level's slab (bit-exact against the oracle on a slice)
    idx = nv.ops.grid_indices(x[:4096], spec)
    assert torch.equal(idx.cpu(), O.hash_indices(x[:4096].cpu(), sc, log2))
    lvl = torch.arange(L, device=DEV)[None, :, None]
    assert bool(((idx >> log2) == lvl).all())
    # linearity in the table: encode(a T1 + b T2) = a encode(T1) + b encode(T2)
    t1 = torch.randn(L << log2, 2, device=DEV, generator=g)
    t2 = torch.randn(L << log2, 2, device=DEV, generator=g)
    y = nv.ops.grid_forward(x, 0.75 * t1 - 1.5 * t2, spec)
    y12 = 0.75 * nv.ops.grid_forward(x, t1, spec) - 1.5 * nv.ops.grid_forward(x, t2, spec)
    assert float((y - y12).abs().max()) < 2e-5
    del t1, t2, y, y12
    # scatter: the eight trilinear weights of a sample sum to one, so every (level, feature) column keeps its gradient mass
    dy = torch.randn(n, 2 * L, device=DEV, generator=g)
    want = dy.double().sum(0).view(L, 2)
    for tmf in (False, True):
        dt = nv.ops.grid_backward(x, _rows_to_tmf(dy) if tmf else dy, spec, tmf=tmf)
        got = dt.view(L, 1 << log2, 2).double().sum(1)
        assert float((got - want).abs().max()) < 1e-2 * float(dy.abs().sum(0).max()) * 1e-3 + 1e-2, tmf
        assert int((dt != 0).any(1).sum()) <= n * L * 8


def test_config3_batch_65536_step_properties():
    import nerf_vo_b200 as nv
    from nerf_vo_b200.synthetic import synthetic_jitters, synthetic_rays
    from nerf_vo_b200.trainer import MappingTrainer

    B, K = 65536, 64
    torch.manual_seed(0)
    model = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(), num_train_data=K).to(DEV)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "hash_table" in name:
                p.normal_(0, 0.1)
    rays, targets = synthetic_rays(B, num_images=K, seed=21)
    jit = synthetic_jitters(B, seed=22)
    dev = lambda d: {k: v.to(DEV) for k, v in d.items()}
    model.train()
    rb = nv.RayBundle(origins=rays["origins"].to(DEV), directions=rays["directions"].to(DEV), pixel_area=rays["pixel_area"].to(DEV),
                      camera_indices=rays["camera_indices"].to(DEV), metadata={"directions_norm": rays["directions_norm"].to(DEV)})
    with torch.no_grad():
        out = model(rb, [j.to(DEV) for j in jit])
    for w, rs, S in zip(out["weights_list"], out["ray_samples_list"], (256, 96, 48)):
        w = w.reshape(B, S)
        assert bool(torch.isfinite(w).all()) and float(w.min()) >= 0.0
        assert float(w.sum(1).max()) <= 1.0 + 1e-4  # alpha compositing: a sub-probability distribution along the ray
        sd = rs.sdist()
        assert sd.shape == (B, S + 1) and bool((sd[:, 1:] >= sd[:, :-1]).all())  # sorted bins (PDF resampling keeps the order)
        assert float(sd.min()) >= 0.0 and float(sd.max()) <= 1.0
    w = out["weights_list"][-1].reshape(B, 48)
    assert float((out["accumulation"].reshape(B) - w.sum(1)).abs().max()) < 1e-4
    assert float(out["rgb"].min()) >= 0.0 and float(out["rgb"].max()) <= 1.0
    n = out["normals"]
    assert bool(torch.isfinite(n).all())
    # two optimizer steps at this batch size: finite, and the parameters of both groups move
    tr = MappingTrainer(model, num_rays=B, use_cuda_graph=False)
    before = [tr.flat[off:off + 1024].clone() for _, off, _ in tr.groups]
    tr.set_inputs(dev(rays), dev(targets), [j.to(DEV) for j in jit])
    losses = [float(tr.train_step()) for _ in range(2)]
    assert all(np.isfinite(losses))
    assert [int(c) for c in tr.step_counts] == [2, 2]
    assert all(not torch.equal(b, tr.flat[off:off + 1024]) for b, (_, off, _) in zip(before, tr.groups))
