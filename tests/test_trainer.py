"""MappingTrainer (nerf-vo_b200/trainer.py) host logic on the GPU.  The numerical trajectory is judged against the CPU oracle stepped by
torch.optim.Adam in tests/test_full_size_parity.py::test_trainer_trajectory_vs_oracle_adam; here: the reference's proposal update schedule
(NS/model_components/ray_samplers.py:596-610) as seen from outside, capture() being free of side effects, the anneal schedule reaching
the CUDA graph (NS/models/nerfacto.py:256-278), and keyframes inserted after capture() being sampled by graph replays
(nerf_vo/mapping/nerfstudio_utils.py:203-241,295-300)."""
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


@pytest.mark.parametrize("graph", [False, True])
def test_trainer_schedule_and_capture_side_effects(graph):
    import nerf_vo_b200 as nv
    from nerf_vo_b200.trainer import MappingTrainer

    rays, targets, jit = _inputs()
    model = _small_model(nv).to(DEV)
    tr = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=graph, proposal_update="reference")
    start = {k: v.detach().clone() for k, v in model.state_dict().items()}
    tr.capture(warmup=2)
    # capture() warms up with real optimizer steps: it must hand back the state it found (parameters, moments, counters, schedule)
    for k, v in model.state_dict().items():
        assert torch.equal(v, start[k]), f"capture() moved {k}"
    assert all(float(t.abs().max()) == 0.0 for t in tr._moment_tensors())
    assert [int(c) for c in tr.step_counts] == [0, 0] and tr.iteration == 0 and tr._ssu == 0
    tr.set_inputs({k: v.to(DEV) for k, v in rays.items()}, {k: v.to(DEV) for k, v in targets.items()}, [j.to(DEV) for j in jit])
    # iterations 0..10 update the proposal networks (step < 10 as the sampler sees it), then every second one (schedule value 1)
    want = [True] * 11 + [False, True, False, True, False]
    prop_keys = [k for k in start if k.startswith("proposal_networks.") and start[k].dtype.is_floating_point and start[k].ndim > 0]
    anneals = []
    for it in range(STEPS):
        before = {k: model.state_dict()[k].detach().clone() for k in prop_keys}
        tr.train_step()
        anneals.append(float(tr._anneal_dev))
        frozen = all(torch.equal(before[k], model.state_dict()[k]) for k in prop_keys)
        assert frozen == (not want[it]), f"iteration {it}: proposal networks {'did not move' if frozen else 'moved'}"
    torch.cuda.synchronize()
    assert [int(c) for c in tr.step_counts] == [STEPS, sum(want)]
    # the device scalar the resampling kernel reads followed the reference's anneal schedule
    assert anneals == pytest.approx([O.anneal_value(it) for it in range(STEPS)], rel=1e-6)


def test_anneal_reaches_the_graph():
    """Same inputs, same parameters, two iterations numbers: anneal 0 (iteration 0: uniform resampling) and ~0.92 (iteration 500) must give
    different losses from the SAME captured graph, each equal to the eager model evaluated with that anneal value."""
    import nerf_vo_b200 as nv
    from nerf_vo_b200.trainer import MappingTrainer

    rays, targets, jit = _inputs()
    model = _small_model(nv).to(DEV)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "hash_table" in n:
                p.normal_(0, 0.3)  # non-uniform proposal weights, so the anneal exponent matters
    tr = MappingTrainer(model, num_rays=B, lr=0.0, eps=1e-15, use_cuda_graph=True)  # lr 0: parameters stay put
    tr.capture(warmup=1)
    tr.set_inputs({k: v.to(DEV) for k, v in rays.items()}, {k: v.to(DEV) for k, v in targets.items()}, [j.to(DEV) for j in jit])
    losses = {}
    for it in (0, 500, 5000):
        tr.iteration = it
        losses[it] = float(tr.train_step())
    assert abs(losses[0] - losses[500]) > 1e-6 * abs(losses[0]) and abs(losses[500] - losses[5000]) > 1e-7 * abs(losses[0]), losses
    rb = nv.RayBundle(origins=rays["origins"].to(DEV), directions=rays["directions"].to(DEV), pixel_area=rays["pixel_area"].to(DEV),
                      camera_indices=rays["camera_indices"].to(DEV), metadata={"directions_norm": rays["directions_norm"].to(DEV)})
    batch = {"image": targets["rgb"].to(DEV), "depth_image": targets["depth"].to(DEV), "normal_image": targets["normal"].to(DEV)}
    for it in (0, 500, 5000):
        model.before_train_iteration(it)
        model.proposal_sampler._steps_since_update = 10 ** 6
        _, ld, _ = model.get_train_loss_dict(rb, batch, [j.to(DEV) for j in jit])
        assert abs(float(sum(ld.values())) - losses[it]) <= 1e-5 * abs(losses[it]), (it, float(sum(ld.values())), losses[it])


def test_graph_replay_samples_keyframes_inserted_after_capture():
    """num_active_frames used to be a by-value kernel argument frozen at capture time."""
    import nerf_vo_b200 as nv
    from nerf_vo_b200.data import DynamicDataManager, DynamicDataManagerConfig
    from nerf_vo_b200.synthetic import synthetic_keyframes
    from nerf_vo_b200.trainer import MappingTrainer

    model = _small_model(nv).to(DEV)
    dm = DynamicDataManager(DynamicDataManagerConfig(train_num_rays_per_batch=B, num_frames=K, frame_height=24, frame_width=32), device=torch.device(DEV))
    synthetic_keyframes(dm.train_dataset, seed=1)
    dm.train_dataset.num_active_frames = 2  # the mapping thread has received two keyframes so far
    tr = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=True, datamanager=dm)
    tr.capture(warmup=1)
    seen = set()
    for _ in range(4):
        tr.train_step()
        seen |= set(dm._last_camera_indices.reshape(-1).tolist())
    assert seen == {0, 1}, seen
    dm.train_dataset.num_active_frames = K  # six more keyframes arrive; no re-capture
    seen = set()
    for _ in range(8):
        tr.train_step()
        seen |= set(dm._last_camera_indices.reshape(-1).tolist())
    assert seen == set(range(K)), seen


@pytest.mark.parametrize("graph", [False, True])
def test_deferred_fields_update_matches_undeferred(graph):
    """defer_fields_update=True runs the fields group's Adam of step k at the start of step k+1 (next to the proposal sampling): after flush()
    parameters, moments and step counters equal the undeferred trainer's after the same steps (both walk the same kernels; the hash-table
    scatter's atomic order is the only difference between two runs), the loss trajectory is the same, and without flush() the fields
    group lags by exactly one update."""
    import nerf_vo_b200 as nv
    from nerf_vo_b200.trainer import MappingTrainer

    rays, targets, jit = _inputs()
    out = []
    for defer in (False, True):
        model = _small_model(nv).to(DEV)
        tr = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=graph, proposal_update="reference", defer_fields_update=defer)
        assert tr.defer_fields == defer
        tr.capture(warmup=2)
        tr.set_inputs({k: v.to(DEV) for k, v in rays.items()}, {k: v.to(DEV) for k, v in targets.items()}, [j.to(DEV) for j in jit])
        losses = [float(tr.train_step()) for _ in range(STEPS)]
        if defer:
            assert [int(c) for c in tr.step_counts] == [STEPS - 1, 13]  # the last fields update is pending
            tr.flush()
            tr.flush()  # idempotent
        torch.cuda.synchronize()
        assert [int(c) for c in tr.step_counts] == [STEPS, 13]
        out.append((losses, tr.flat.detach().clone(), [m.detach().clone() for m in tr._moment_tensors()]))
    (l0, p0, m0), (l1, p1, m1) = out
    assert l1 == pytest.approx(l0, rel=1e-3), (l0, l1)
    init = torch.cat([torch.cat([p.detach().reshape(-1), p.new_zeros((-p.numel()) % 4)]) for p in _small_model(nv).to(DEV).parameters() if p.requires_grad])
    d0, d1 = (p0 - init).double(), (p1 - init).double()  # what 16 steps moved (Adam at lr 1e-2 amplifies the scatter's atomic-order noise on
    cos = float((d0 * d1).sum() / (d0.norm() * d1.norm()))  # entries with tiny gradients: judged by the movement as a whole, as the oracle test does)
    assert cos > 0.995, cos
    assert abs(float(d0.norm()) - float(d1.norm())) < 2e-2 * float(d0.norm())
    for a, b in zip(m0, m1):
        assert float((a - b).norm()) <= 1e-1 * float(a.norm()) + 1e-12  # same chaos-amplified noise, seen through the moments
