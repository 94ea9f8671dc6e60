"""Oracle restatement vs the UNMODIFIED reference executed live (build container only;
skipped on the GPU box where /root/reference does not exist)."""
import os
import warnings

import pytest
import torch

import nerfacto_oracle as O
import reference_harness as rh

pytestmark = pytest.mark.skipif(not rh.available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("seed", [0, 7])
def test_full_step_matches_reference(seed):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = rh.build_reference_model(main_log2=11, prop_log2=9, num_images=6, seed=seed)
        torch.manual_seed(seed + 1)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "hash_table" in n:
                    p.normal_(0, 0.2)
        sd = rh.reference_state(m)
        B = 48
        rays, targets = O.synthetic_rays(B, num_images=6, seed=seed + 5)
        out, ld, jit, _ = rh.reference_step(m, rays, targets, step=100 * seed)
    cfg = O.ModelCfg(main_grid=O.GridCfg(log2_hashmap_size=11), prop_grids=(O.GridCfg(5, 16, 128, 9), O.GridCfg(5, 16, 256, 9)), num_images=6)
    P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    o2, L2, _ = O.mapping_step(P, cfg, rays, targets, jit, anneal=O.anneal_value(100 * seed))
    for k in ("rgb", "accumulation", "depth", "expected_depth", "normals", "pred_normals", "prop_depth_0", "prop_depth_1"):
        torch.testing.assert_close(o2[k], out[k], rtol=0, atol=1e-6, msg=k)
    for k, v in ld.items():
        torch.testing.assert_close(L2[k].detach().reshape(()), v.detach().reshape(()), rtol=1e-5, atol=1e-12, msg=k)
    gref = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
    for k in P:
        if k in gref:
            torch.testing.assert_close(P[k].grad, gref[k], rtol=1e-4, atol=1e-9, msg=k)
