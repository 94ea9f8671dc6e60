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


def test_proposal_update_schedule_matches_reference():
    """The `updated` predicate of the reference's own ProposalNetworkSampler (NS/model_components/ray_samplers.py:591-612, schedule
    NS/models/nerfacto.py:202-207 with NeRF-VO's update_every = 5, warm-up 5000), driven for 8192 iterations the way nerfstudio's trainer
    drives it (forward decides, AFTER_TRAIN_ITERATION callback counts), against MappingTrainer's host-side decision."""
    import nerf_vo_b200 as nv
    from nerf_vo_b200.trainer import MappingTrainer

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = rh.build_reference_model(main_log2=9, prop_log2=8, num_images=4, seed=0)
    ps = m.proposal_sampler
    ref = []
    for it in range(8192):
        upd = bool(ps._steps_since_update > ps.update_sched(ps._step) or ps._step < 10)  # ray_samplers.py:596
        if upd:
            ps._steps_since_update = 0  # :611-612
        ref.append(upd)
        ps.step_cb(it)  # :591-594
    cfg = nv.NerfactoModelConfig(log2_hashmap_size=9)
    for a in cfg.proposal_net_args_list:
        a["log2_hashmap_size"] = 8
    tr = MappingTrainer(nv.ExtendedNerfactoModel(cfg, num_train_data=4), num_rays=16, use_cuda_graph=False, device=torch.device("cpu"),
                        proposal_update="reference")
    tr._ssu = 0
    ours = []
    for it in range(8192):
        tr.iteration = it
        upd = tr._updated_now()
        ours.append(upd)
        tr._ssu = 1 if upd else tr._ssu + 1
    assert ours == ref


def test_reference_loads_our_checkpoint_strictly(tmp_path):
    """A checkpoint written by nerf_vo_b200.checkpoint.save_checkpoint goes through the reference's own strict Module.load_state_dict
    (what VanillaPipeline.load_state_dict tries first, NS/pipelines/base_pipeline.py:127-128) after the reference's prefix stripping."""
    import nerf_vo_b200 as nv

    cfg = nv.NerfactoModelConfig(log2_hashmap_size=9)
    for a in cfg.proposal_net_args_list:
        a["log2_hashmap_size"] = 8
    torch.manual_seed(3)
    ours = nv.ExtendedNerfactoModel(cfg, num_train_data=4)
    path = os.path.join(str(tmp_path), nv.checkpoint.checkpoint_name(42))
    nv.checkpoint.save_checkpoint(path, 42, ours)
    loaded = torch.load(path, map_location="cpu", weights_only=False)
    assert loaded["step"] == 42
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = rh.build_reference_model(main_log2=9, prop_log2=8, num_images=4, seed=1)
    state = {k[len("_model."):]: v for k, v in loaded["pipeline"].items() if k.startswith("_model.")}  # base_pipeline.py:113-116
    m.load_state_dict(state, strict=True)
    own = ours.state_dict()
    for k, v in m.state_dict().items():
        if k in own:
            assert torch.equal(v, own[k]), k
