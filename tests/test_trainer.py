"""MappingTrainer (nerf-vo_b200/trainer.py) against the step the reference's Trainer.train_iteration performs (NS/engine/trainer.py:455-494):
forward + loss_dict + backward through the public model API, one torch.optim.Adam per parameter group ("fields", "proposal_networks";
NS/engine/optimizers.py:138-150) with zero_grad(set_to_none=True), and ProposalNetworkSampler's own update schedule."""
import numpy as np
import pytest
import torch

import nerfacto_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K, B, STEPS = 8, 256, 16


def _small_model(nv):
    torch.manual_seed(0)
    cfg = nv.NerfactoModelConfig(log2_hashmap_size=14)
    for a in cfg.proposal_net_args_list:
        a["log2_hashmap_size"] = 12
    return nv.ExtendedNerfactoModel(cfg, num_train_data=K)


def _inputs():
    rays, targets = O.synthetic_rays(B, num_images=K, seed=3)
    jit = O.synthetic_jitters(B)
    return rays, targets, jit


def _reference_style_loop(nv, model, rays, targets, jit):
    """What nerfstudio's trainer does, with this repo's modules behind the reference's model API."""
    model = model.to(DEV).train()
    groups = model.get_param_groups()
    opts = [torch.optim.Adam(groups[k], lr=1e-2, eps=1e-15) for k in ("fields", "proposal_networks")]
    rb_kw = dict(origins=rays["origins"].to(DEV), directions=rays["directions"].to(DEV), pixel_area=rays["pixel_area"].to(DEV),
                 camera_indices=rays["camera_indices"].to(DEV))
    batch = {"image": targets["rgb"].to(DEV), "depth_image": targets["depth"].to(DEV), "normal_image": targets["normal"].to(DEV)}
    losses, updated = [], []
    for it in range(STEPS):
        for o in opts:
            o.zero_grad(set_to_none=True)
        rb = nv.RayBundle(metadata={"directions_norm": rays["directions_norm"].to(DEV)}, **rb_kw)
        _, loss_dict, _ = model.get_train_loss_dict(rb, batch, [j.to(DEV) for j in jit])
        updated.append(model.proposal_sampler._steps_since_update == 0)  # reset by the sampler exactly when it let gradients through
        total = sum(loss_dict.values())
        total.backward()
        for o in opts:
            o.step()
        model.after_train_iteration(it)
        losses.append(float(total))
    return losses, updated, {k: v.detach().clone() for k, v in model.state_dict().items()}


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_follows_the_reference_step_and_schedule(graph):
    import nerf_vo_b200 as nv
    from nerf_vo_b200.trainer import MappingTrainer

    rays, targets, jit = _inputs()
    ref_losses, ref_updated, ref_state = _reference_style_loop(nv, _small_model(nv), rays, targets, jit)
    # iterations 0..10 update the proposal networks (step < 10 seen by the sampler), then every second one (schedule value 1)
    assert ref_updated == [True] * 11 + [False, True, False, True, False]

    model = _small_model(nv).to(DEV)
    tr = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=graph, proposal_update="reference")
    start = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr.capture(warmup=1)
    # the warm-up steps trained the model: restore the initial state (parameters, moments, counters) before the compared run
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(start[k])
    for t in (tr.exp_avg, tr.exp_avg_sq):
        t.zero_()
    for c in tr.step_counts:
        c.zero_()
    tr.iteration, tr._ssu = 0, 0
    tr.set_inputs({k: v.to(DEV) for k, v in rays.items()}, {k: v.to(DEV) for k, v in targets.items()}, [j.to(DEV) for j in jit])
    losses, prop_keys = [], [k for k in start if k.startswith("proposal_networks.") and start[k].dtype.is_floating_point and start[k].ndim > 0]
    for it in range(STEPS):
        before = {k: model.state_dict()[k].detach().clone() for k in prop_keys}
        losses.append(float(tr.train_step()))
        frozen = all(torch.equal(before[k], model.state_dict()[k]) for k in prop_keys)
        assert frozen == (not ref_updated[it]), f"iteration {it}: proposal networks {'did not move' if frozen else 'moved'}"
    torch.cuda.synchronize()
    assert [int(c) for c in tr.step_counts] == [STEPS, sum(ref_updated)]
    # same kernels on both sides; what differs is the order of the atomic scatter sums and Adam's fused arithmetic.  With eps = 1e-15 an
    # entry whose gradient is rounding noise moves by +-lr whatever its size, so parameters are compared on the bulk of their entries.
    np.testing.assert_allclose(losses, ref_losses, rtol=2e-2, atol=1e-5)
    now = model.state_dict()
    for k, v in ref_state.items():
        if not v.dtype.is_floating_point or v.ndim == 0:
            continue
        moved = (ref_state[k] - start[k]).abs().max()
        far = ((now[k] - v).abs() > 0.05 * moved + 1e-6).float().mean()
        # measured noise floor (tools/trainer_noise.py, the loop against ITSELF): up to 3 % of the entries of the 64-element head biases
        # land further apart than this, below 1 % elsewhere; a wrong schedule or a racing optimizer moves nearly all of them
        # (tensors of a few elements — output biases — are left to the loss trajectory: one noisy entry of three is already 33 %)
        if v.numel() >= 64:
            assert float(far) < (0.15 if v.numel() >= 256 else 0.3), (k, float(far), float(moved))
