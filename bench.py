#!/usr/bin/env python
"""bench.py — NeRF mapping train rays/s (forward + losses + backward + fused Adam) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--rays R] [--log2-hashmap L]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workloads (BASELINE.json `configs`, 0-based index in brackets; `--config` takes the 1-based number used by SURVEY.md section 8):
  --config 2 [1] (default, the configuration the metric is quoted on): NeRF-VO Replica mapping step — 4096 rays per GPU, proposal sampling
             256/96 + 48 nerf samples, 16 x 2^19 x 2 main hash grid, two 5 x 2^17 x 2 proposal grids, rgb + interlevel + distortion + DS-NeRF
             depth + MonoSDF normal losses, predicted normals on; weak scaling (4096 rays on every GPU);
  --config 3 [2]: ScanNet-config mapping, 65 536 rays per GPU, 512 keyframes of 320 x 240, depth + normal supervision;
  --config 4 [3]: 2^21-row main table, 262 144 rays per step GLOBALLY, sharded over the ranks (strong scaling), gradient exchange at 2/4/8;
  --config 5 [4]: evaluation full-frame render 1200 x 680, forward only, image rows sharded over the ranks.
Synthetic Replica-shaped rays / keyframes, random-init parameters (reference init).

Prints ONE JSON line (rank 0).
  value       rays/s with the batch resident in HBM, CUDA events over K CUDA-graph replays of the whole step;
  e2e         the same metric through the public trainer API as the reference's Trainer.train_iteration window sees it
              (NS/engine/trainer.py:258-266): the step starts from the resident keyframe store — pixel sampling, image / depth / normal
              gather and ray generation run inside the step (fused prologue kernel) — with that step's random draws copied from pinned
              HOST memory (H2D) and the loss read back to the host (D2H) inside the timed region;
  roofline    the dominant kernel of the step, timed live with CUDA events (L2 flushed between launches) against MEASURED_PEAKS.json;
              the other hot kernels under `roofline.others` (hash grid against HBM, field networks against the tensor peak);
  cpu_baseline  the CPU oracle port timed on this box's host cores on the same batch size (N=1 only).
`--impl reference` times the reference's own CPU algorithm (oracle port, validated against the live reference in the build container; the
Python reference cannot travel to the GPU box) with all host threads on the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures of the same kernels / shapes at config 2
# (profiles/): far below the algorithmic bytes where the tables are served from the 126 MB L2.  None = no capture of this kernel at this shape.
RECORDED_TRAFFIC = {"k_grid_fwd_tmh": 56.29e6, "k_grid_bwd_run": 52.2e6, "k_grid_fwd_tmh_jac": 73.8e6}

METRIC = "nerf_mapping_train_rays_per_s"
UNIT = "rays/s"
PRESETS = {
    2: {"rays": 4096, "global_rays": None, "log2": 19, "images": 192, "frame": (360, 640), "scaling": "weak",
        "name": "nerf-vo replica mapping step (BASELINE configs[1])"},
    3: {"rays": 65536, "global_rays": None, "log2": 19, "images": 512, "frame": (240, 320), "scaling": "weak",
        "name": "scannet-config mapping step, depth + normal supervision (BASELINE configs[2])"},
    4: {"rays": None, "global_rays": 262144, "log2": 21, "images": 192, "frame": (360, 640), "scaling": "strong",
        "name": "ray-batch sharded mapping, 2^21 main table, 262144 rays/step globally (BASELINE configs[3])"},
    5: {"rays": None, "global_rays": 816000, "log2": 19, "images": 192, "frame": (680, 1200), "scaling": "strong",
        "name": "evaluation full-frame render 1200x680, forward only (BASELINE configs[4])"},
}


_JSON_OUT = None


def claim_stdout() -> None:
    """stdout carries ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its version banner there at
    communicator creation): fd 1 is pointed at stderr for the lifetime of the process and the JSON line goes to a private duplicate."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload_string(cfg: dict, B: int, world: int) -> str:
    return (f"{cfg['name']}: {B} rays/GPU x {world} GPU, proposal 256/96 + 48 samples, hash 16x2^{cfg['log2']}x2 + 2x(5x2^17x2), "
            "rgb+interlevel+distortion+depth+normal losses")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    """Dense bf16/fp16 tensor peak in TFLOP/s: the burst figure (kernels timed alone)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), "measured burst (MEASURED_PEAKS.json)"
    return 1590.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) >= 6)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_model_name() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.lower().startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's torch path
# ---------------------------------------------------------------------------------------------------------------------
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nerfacto_oracle as O

    return O


def cpu_port_rays_per_s(num_rays: int, steps: int, warmup: int, threads: int, log2: int = 19, images: int = 192):
    """Times the CPU oracle port (oracle/nerfacto_oracle.py) on `num_rays` rays: full forward + losses + backward of one mapping step."""
    import torch

    O = _oracle()
    torch.set_num_threads(threads)
    cfg = O.ModelCfg(main_grid=O.GridCfg(log2_hashmap_size=log2), num_images=images)
    P = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, seed=0).items()}
    times, split = [], {"forward_s": 0.0, "losses_s": 0.0, "backward_s": 0.0}
    for i in range(warmup + steps):
        rays, targets = O.synthetic_rays(num_rays, num_images=images, seed=1234 + i)
        jit = O.synthetic_jitters(num_rays, seed=99 + i)
        for p in P.values():
            p.grad = None
        t0 = time.perf_counter()
        out = O.mapping_forward(P, cfg, rays["origins"], rays["directions"], rays["camera_indices"], jit, 1.0, True, True)
        t1 = time.perf_counter()
        L = O.mapping_losses(cfg, out, targets["rgb"], targets.get("depth"), rays.get("directions_norm"), targets.get("normal"))
        total_loss = sum(L.values())
        t2 = time.perf_counter()
        total_loss.backward()
        t3 = time.perf_counter()
        if i >= warmup:
            times.append(t3 - t0)
            split["forward_s"] += t1 - t0
            split["losses_s"] += t2 - t1
            split["backward_s"] += t3 - t2
    total = sum(times)
    split = {k: v / len(times) for k, v in split.items()}
    return num_rays * len(times) / total, total / len(times), split


def cpu_port_frame_rays_per_s(num_rays: int, steps: int, warmup: int, threads: int):
    """Evaluation (config 5) on the CPU oracle: eval-mode forward of `num_rays` rays per step in the reference's 4096-ray chunks."""
    import torch

    O = _oracle()
    torch.set_num_threads(threads)
    cfg = O.ModelCfg()
    P = O.init_params(cfg, seed=0, table_std=0.1)
    times = []
    for i in range(warmup + steps):
        rays, _ = O.synthetic_rays(num_rays, seed=77 + i)
        t0 = time.perf_counter()
        with torch.no_grad():
            for a in range(0, num_rays, 4096):
                O.mapping_forward(P, cfg, rays["origins"][a:a + 4096], rays["directions"][a:a + 4096], rays["camera_indices"][a:a + 4096], None, 1.0, False, False)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return num_rays * len(times) / sum(times), sum(times) / len(times)


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_total = args.steps + args.warmup
    if args.config == 5:
        rays = 16384 if n_total <= 40 else 4096
        rps, sec = cpu_port_frame_rays_per_s(rays, args.steps, args.warmup, threads)
        split = None
        sample = f"{rays} rays of the 816000-ray frame per step (eval-mode forward in 4096-ray chunks)"
        metric = "nerf_eval_render_rays_per_s"
    else:
        # the full 4096-ray batch of config 2 (~1.2 s per step on 16 cores); a bounded 4096-ray sample of the larger configs' batches
        rays = 4096 if n_total <= 120 else 1024
        rps, sec, split = cpu_port_rays_per_s(rays, args.steps, args.warmup, threads, cfg["log2"], cfg["images"])
        full = cfg["rays"] or cfg["global_rays"]
        sample = (f"the full {rays}-ray batch per step" if rays == full else f"{rays} rays per step (bounded sample of the {full}-ray batch)") + ", full-size tables"
        metric = METRIC
    line = {
        "metric": metric, "value": rps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": cfg["name"], "sample": sample},
        "cpu_baseline": {"value": rps, "unit": UNIT, "cores": threads, "cpu_model": cpu_model_name(), "kind": "port", "phases_s_per_step": split,
                         "sample": f"{args.steps} timed steps ({args.warmup} warm-up) x {sample}; the reference's torch algorithm as restated by the oracle port "
                                   "(pinned to the live reference in the build container; the Python reference is not on this box)"},
        "e2e": {"value": rps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
def _dist_setup():
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the nvo_b200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev


def _timed_us(torch, flush, fn, reps=20):
    """Average device time of fn() in microseconds: CUDA events on the launching stream, L2 flushed (256 MB write) before every launch."""
    ev = []
    for i in range(3 + reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev[3:]) / reps * 1e3


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    import nerf_vo_b200 as nv
    from nerf_vo_b200.synthetic import synthetic_jitters, synthetic_rays
    from nerf_vo_b200.trainer import MappingTrainer

    rank, world, local, dev = _dist_setup()
    nv._lib.load()
    K_IMG = cfg["images"]
    B = args.rays or cfg["rays"] or cfg["global_rays"] // world
    log2 = args.log2_hashmap or cfg["log2"]

    torch.manual_seed(0)  # (the trainer broadcasts rank 0's parameters anyway; rays differ per rank)
    # --pose-opt: the model's camera optimizer on, as in the reference's mapping step (NS/models/nerfacto.py:130 default "SO3xR3"; group
    # "camera_opt").  The headline stays the configuration of BASELINE.json / round 1 (and of the CPU reference arm, which does not optimise
    # poses either); the `pose_opt` leg below reports the step with it
    pose_mode = "SO3xR3" if args.pose_opt else "off"
    model = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(log2_hashmap_size=log2, camera_optimizer_mode=pose_mode), num_train_data=K_IMG).to(dev)
    defer = args.defer_fields == "on" or (args.defer_fields == "auto" and world > 1)
    trainer = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=not args.no_graph, exchange=args.exchange, defer_fields_update=defer)
    trainer.iteration = 2000  # past the proposal-weight anneal window (first 1000 of NeRF-VO's 8192 iterations): steady-state step

    n_pool = 8 if B <= 65536 else 3
    host_batches, dev_batches = [], []
    for i in range(n_pool):
        rays, targets = synthetic_rays(B, num_images=K_IMG, seed=1234 + 1000 * rank + i)
        jit = synthetic_jitters(B, seed=99 + 1000 * rank + i)
        host_batches.append(({k: v.pin_memory() for k, v in rays.items()}, {k: v.pin_memory() for k, v in targets.items()}, [j.pin_memory() for j in jit]))
        dev_batches.append(({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in jit]))
    trainer.set_inputs(*dev_batches[0])
    trainer.capture(warmup=3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(feed, read_loss: bool, steps: int):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h2d = 0
        barrier()
        ev0.record()
        for s in range(steps):
            h2d = feed(s)
            loss = trainer.train_step()
            if read_loss:
                loss_host = float(loss)  # D2H of the step's result
        trainer.flush()  # defer_fields_update: the last step's pending fields-group update belongs to the timed steps
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, h2d

    # ---- value: inputs resident in HBM (one device-to-device copy of the packed batch per step) -----------------------------------------
    packed_dev = [tuple(t.to(dev) for t in trainer.pack_host_batch(*hb)) for hb in host_batches]
    def settle(step_fn):
        """Warm-up: max(3, W) untimed steps, extended to ~150 ms of stepping (a count fixed by the batch size, identical on every rank: the
        steps contain collectives).  A 4096-ray step is under a millisecond, so W steps alone end before the clocks of all ranks have ramped up
        and the ranks' launch queues have filled — the first timed steps of a short window then carry that ramp (r02: 8 GPUs, 40 timed
        steps read 0.96 ms per step, 100 steps 0.87 ms)."""
        n = max(3, args.warmup, min(200, int(0.15 / (B * 2e-7))))
        for s in range(n):
            step_fn(s)
            if s % 16 == 15:
                torch.cuda.synchronize()
        return n

    def warm_value(s):
        trainer.set_inputs_packed(packed_dev[s % n_pool])
        trainer.train_step()

    settle(warm_value)
    with ClockSampler(local) as clk:
        ms, _ = timed(lambda s: trainer.set_inputs_packed(packed_dev[s % n_pool]), False, args.steps)
    launches_per_step = trainer.launches_per_step
    # ---- secondary end-to-end leg: ray batches packed on the HOST (pinned), two H2D copies per step, loss read back -------------------------
    packed_batches = [trainer.pack_host_batch(*hb) for hb in host_batches]
    for s in range(3):
        trainer.set_inputs_packed(packed_batches[s % n_pool])
        float(trainer.train_step())
    ms_hb, h2d_hb = timed(lambda s: trainer.set_inputs_packed(packed_batches[s % n_pool]), True, args.steps)
    final_loss = float(trainer.loss)
    rays_per_s = world * B * args.steps / (ms * 1e-3)
    host_batch_fed = {"value": world * B * args.steps / (ms_hb * 1e-3), "unit": UNIT, "ms_per_step": ms_hb / args.steps, "h2d_bytes_per_step": h2d_hb,
                      "d2h_bytes_per_step": 4, "note": "ray batch (origins, directions, targets, jitters) packed on the host; no pixel sampling / ray generation in the step"}

    # ---- N > 1: the exchange step alone ---------------------------------------------------------------------------------------------------
    exchange = None
    if world > 1:
        reps = 20
        for _ in range(3):
            trainer._exchange(); trainer._optimizer()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            trainer._exchange(); trainer._optimizer()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cs = torch.stack([trainer.flat.double().sum(), trainer.flat.double().abs().sum()])
        allcs = [torch.zeros_like(cs) for _ in range(world)]
        dist.all_gather(allcs, cs)
        consistent = all(torch.equal(allcs[0], c) for c in allcs)  # every replica holds bit-identical parameters after the run
        nbytes = trainer.flat.numel() * 4
        link = (world - 1) / world * nbytes
        exchange = {"arm": trainer.exchange, "us_per_step": float(t) * 1e3, "flat_bytes": nbytes, "nvlink_bytes_per_direction_per_gpu": link,
                    "achieved_GBs_per_direction": link / (float(t) * 1e-3) / 1e9, "peer_copy_peak_GBs": 770.0,
                    "frac_of_peer_copy_peak": link / (float(t) * 1e-3) / 1e9 / 770.0,
                    "replicas_bit_identical": bool(consistent), "error_word": trainer.peer.error_word() if trainer.peer is not None else 0}

    # ---- roofline (rank 0): the step's hot kernels timed alone, L2 flushed between launches ---------------------------------------------
    roofline = cpu = None
    if rank == 0 and not args.no_roofline:
        roofline = measure_roofline(torch, nv, model, trainer, dev, B)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_cpu = min(B, 4096)
        v, sec, split = cpu_port_rays_per_s(n_cpu, 5, 1, threads, log2, K_IMG)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "cpu_model": cpu_model_name(), "kind": "port", "phases_s_per_step": split,
               "sample": f"5 timed steps (1 warm-up) x {n_cpu} rays" + (" (the full batch)" if n_cpu == B else f" (bounded sample of the {B}-ray batch)") +
                         ", full-size tables, oracle port of the reference torch path"}

    # ---- e2e: the step as Trainer.train_iteration runs it — fed by the resident keyframe store through the fused prologue kernel (pixel
    # sampling + gather + ray generation inside the CUDA graph), the step's random draws arriving from pinned host memory, loss read back ----
    from nerf_vo_b200.data import DynamicDataManager, DynamicDataManagerConfig
    from nerf_vo_b200.synthetic import synthetic_keyframes

    fh, fw = cfg["frame"]
    dm = DynamicDataManager(DynamicDataManagerConfig(train_num_rays_per_batch=B, num_frames=K_IMG, frame_height=fh, frame_width=fw), device=dev)
    synthetic_keyframes(dm.train_dataset, seed=4321 + rank)
    trainer.datamanager, trainer.external_draws = dm, True
    trainer.capture(warmup=3)
    g = torch.Generator().manual_seed(555 + rank)
    draws = [trainer.pack_host_draws(torch.rand(B, 3, generator=g), [torch.rand(B, 1, generator=g) for _ in range(3)]) for _ in range(n_pool)]
    def warm_e2e(s):
        trainer.set_draws_packed(draws[s % n_pool])
        float(trainer.train_step())

    settle(warm_e2e)
    ms_e2e, h2d = timed(lambda s: trainer.set_draws_packed(draws[s % n_pool]), True, args.steps)
    e2e_rays_per_s = world * B * args.steps / (ms_e2e * 1e-3)
    e2e_loss = float(trainer.loss)
    prologue_us = None
    if rank == 0 and not args.no_roofline:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        prologue_us = _timed_us(torch, flush, lambda: dm.next_train(0, u=trainer.inputs["u"]))
        del flush
    resident = int(sum(t.numel() * 4 for t in (dm.train_dataset.frames_color, dm.train_dataset.frames_depth, dm.train_dataset.frames_normal)))
    trainer.datamanager, trainer.external_draws = None, False

    # ---- the reference's own proposal update schedule (NS/model_components/ray_samplers.py:596-610 with NeRF-VO's update_every = 5, warm-up
    # 5000 of 8192 iterations): past the warm-up only every 6th step sends gradients to the proposal networks. `value` is the every-step case.
    ref_sched = None
    if world == 1 and not args.no_schedule_leg:
        trainer2 = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=not args.no_graph, proposal_update="reference")
        trainer2.set_inputs(*dev_batches[0])
        trainer2.capture(warmup=3)
        trainer2.iteration, trainer2._ssu = 6000, 1
        for s_ in range(12):
            trainer2.set_inputs(*dev_batches[s_ % n_pool])
            trainer2.train_step()
        k = (args.steps + 5) // 6 * 6
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0 = int(trainer2.step_counts[1]) if len(trainer2.step_counts) > 1 else 0
        e0.record()
        for s_ in range(k):
            trainer2.set_inputs(*dev_batches[s_ % n_pool])
            trainer2.train_step()
        e1.record()
        torch.cuda.synchronize()
        ms_rs = e0.elapsed_time(e1)
        ref_sched = {"value": B * k / (ms_rs * 1e-3), "unit": UNIT, "ms_per_step": ms_rs / k, "steps": k,
                     "proposal_updates": (int(trainer2.step_counts[1]) - c0) if len(trainer2.step_counts) > 1 else None,
                     "note": "iterations 6000.. of NeRF-VO's 8192: proposal networks receive gradients every 6th step (reference schedule); inputs resident"}
        del trainer2

    # ---- the same step with the camera optimizer on (SO3xR3): ray gradients of all three sampling levels -> pose deltas -> "camera_opt" Adam ----
    pose_leg = None
    if world == 1 and not args.no_schedule_leg and not args.pose_opt and log2 <= 19:
        torch.manual_seed(0)
        model_p = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(log2_hashmap_size=log2, camera_optimizer_mode="SO3xR3"), num_train_data=K_IMG).to(dev)
        trainer_p = MappingTrainer(model_p, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=not args.no_graph)
        trainer_p.iteration = 2000
        trainer_p.set_inputs(*dev_batches[0])
        trainer_p.capture(warmup=3)
        for s_ in range(max(3, args.warmup)):
            trainer_p.set_inputs_packed(packed_dev[s_ % n_pool])
            trainer_p.train_step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s_ in range(args.steps):
            trainer_p.set_inputs_packed(packed_dev[s_ % n_pool])
            trainer_p.train_step()
        e1.record()
        torch.cuda.synchronize()
        ms_p = e0.elapsed_time(e1)
        pose_leg = {"value": B * args.steps / (ms_p * 1e-3), "unit": UNIT, "ms_per_step": ms_p / args.steps, "launches_per_step": trainer_p.launches_per_step,
                    "pose_moved_max_abs": float(model_p.camera_optimizer.pose_adjustment.detach().abs().max()),
                    "note": "camera optimizer on (SO3xR3, NS/models/nerfacto.py:130): d loss / d pose through every level's sample positions and the "
                            "direction encoding, camera_opt Adam (1e-4 -> 1e-5 exponential); inputs resident"}
        del trainer_p, model_p

    if rank == 0:
        table_mb = (16 << log2) * 2 * 4 / 2**20
        line = {
            "metric": METRIC, "value": rays_per_s, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_string(cfg, B, world), "baseline_config": args.config,
                       "precision": "fp16 tensor-core operands (hash features, field MLPs) with fp32 accumulation; fp32 tables, proposal networks, per-ray ops, optimizer",
                       "rays_per_gpu": B, "global_rays": world * B, "parallelism": f"dp{world} (ray sharding; gradient exchange: {trainer.exchange})",
                       "step": "zero-grad + forward + losses + backward + fused Adam, proposal networks updated every step, camera poses optimised ("
                               + pose_mode + ": ray gradients of all three levels -> pose deltas, Adam under the exponential schedule)",
                       "l2": f"no explicit flush: main table + gradient + Adam moments = {4 * table_mb:.0f} MiB streamed per step exceed the 126 MB L2",
                       "cuda_graph": not args.no_graph,
                       "defer_fields_update": bool(trainer.defer_fields),
                       "defer_note": ("the fields group's exchange + Adam of step k is launched at the start of step k+1 next to its proposal sampling; "
                                      "the timed region ends with flush(), so K steps contain K fields updates") if trainer.defer_fields else None},
            "e2e": {"value": e2e_rays_per_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "path": "MappingTrainer(datamanager=DynamicDataManager, external_draws=True): set_draws_packed (pinned host -> device) + train_step "
                            "(prologue kernel: pixel sampling + rgb/depth/normal gather + ray generation from the resident keyframe store; forward; losses; "
                            "backward; Adam) + float(loss)",
                    "keyframes": K_IMG, "frame": f"{fh}x{fw}", "resident_bytes": resident, "prologue_us": prologue_us, "final_loss": e2e_loss},
            "gpu_launches": int(launches_per_step * args.steps),
            "gpu_launches_per_step": int(launches_per_step),
            "clocks": clk.summary(),
            "roofline": roofline,
            "cpu_baseline": cpu,
            "exchange": exchange,
            "host_batch_fed": host_batch_fed,
            "reference_schedule": ref_sched, "pose_opt": pose_leg,
            "final_loss": final_loss,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_roofline(torch, nv, model, trainer, dev, B):
    """The step's hot kernels, each repeated alone on the LAST TIMED STEP's own sample positions with the L2 flushed between launches.
    Headline = the hash-table scatter (largest share of the step in the ncu launch lists under profiles/), three launches per step."""
    peak, peak_src = load_peaks()
    tpeak, tpeak_src = load_tensor_peak()
    N = B * 48
    enc = model.field.mlp_base.encoder
    x = model.field._cache["x"].detach().clone()
    assert x.shape == (N, 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    table = enc.hash_table.detach()
    timed_us = lambda fn, reps=20: _timed_us(torch, flush, fn, reps)

    # sample positions of the two proposal levels: the sampler run once more on the last batch (no gradients)
    with torch.no_grad():
        model.proposal_sampler._steps_since_update = 0
        bundle = model.set_nears_and_fars(trainer._bundle())
        _, _, rs_list = model.proposal_sampler(bundle, density_fns=model.density_fns, jitters=[trainer.inputs[f"jitter{k}"] for k in range(3)])
        xs = [nv.ops.contract_normalize(rs.frustums.get_positions().reshape(-1, 3).contiguous())[0] for rs in rs_list[:2]]
    torch.cuda.synchronize()
    launches = [(f"main grid 16 x 2^{enc.spec.log2_T}", x, enc.spec, torch.zeros_like(table))]
    for i, pn in enumerate(model.proposal_networks):
        launches.append((f"proposal grid {i} 5 x 2^{pn.encoding.spec.log2_T}", xs[i], pn.encoding.spec, torch.zeros_like(pn.encoding.hash_table.detach())))
    parts, tot_bytes, tot_us = [], 0, 0.0
    for name, xl, spec, scratch in launches:
        n_l, L_l = xl.shape[0], spec.n_levels
        dy_l = torch.randn(nv.ops.tmh_numel(n_l, spec.out_dim), device=dev)
        us = timed_us(lambda: nv.ops.grid_backward(xl, dy_l, spec, dtable=scratch, tmf=True))
        b_s = 12 + L_l * 2 * 4 + L_l * 8 * 2 * 4  # xyz + dL/dy (2L fp32) + L levels x 8 corners x 2 fp32 scattered once
        b_16 = 12 + L_l * 2 * 2 + L_l * 8 * 2 * 2  # the same on SURVEY 8d's fp16 byte accounting
        parts.append({"launch": name, "samples": n_l, "launch_us": us, "algorithmic_bytes_per_sample": b_s, "achieved_GBs": b_s * n_l / us / 1e3,
                      "frac": b_s * n_l / us / 1e3 / peak, "frac_fp16_accounting": b_16 * n_l / us / 1e3 / peak})
        tot_bytes += b_s * n_l
        tot_us += us
        del dy_l
    ach = tot_bytes / tot_us / 1e3
    roofline = {"bound": "hbm", "kernel": "hash-table scatter (fp32 red.global.add; 3 launches per step: main 16-level grid + two 5-level proposal grids, "
                                           "L2 flushed between launches)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": RECORDED_TRAFFIC.get("k_grid_bwd_run") if B == 4096 else None,
                "peak_source": peak_src, "launch_us": tot_us / len(parts), "algorithmic_bytes_per_launch": tot_bytes / len(parts),
                "launches_per_step": len(parts), "per_launch": parts, "inputs": "sample positions of the last timed step (all three levels)"}
    del launches
    extras = []
    us = timed_us(lambda: nv.ops.grid_forward_jac(x, table, enc.spec))
    b_s = 12 + 16 * 8 * 2 * 4 + 16 * 2 * 2 + 3 * 16 * 2 * 2  # xyz + 16 x 8 corner rows (fp32 pairs) + 32 fp16 features + 96 fp16 derivatives
    b_16 = 12 + 16 * 8 * 2 * 2 + 16 * 2 * 2 + 3 * 16 * 2 * 2
    extras.append({"kernel": "main hash grid forward (table -> fp16 TMH feature tiles + saved d feature/dx)", "bound": "hbm",
                   "launch_us": us, "achieved": b_s * N / us / 1e3, "peak": peak, "unit": "GB/s", "frac": b_s * N / us / 1e3 / peak,
                   "frac_fp16_accounting": b_16 * N / us / 1e3 / peak,
                   "algorithmic_bytes_per_sample": b_s, "traffic": RECORDED_TRAFFIC.get("k_grid_fwd_tmh_jac") if B == 4096 else None})
    n_par = trainer.groups[0][2]
    p2, m2, v2 = trainer.flat[:n_par].clone(), torch.zeros(n_par, device=dev), torch.zeros(n_par, device=dev)  # copies: the trainer's state is not touched
    g2, cnt = trainer.grad[:n_par].clone(), torch.zeros(1, dtype=torch.int32, device=dev)
    us = timed_us(lambda: nv.ops.adam_step(p2, g2, m2, v2, cnt, 1e-2, 0.9, 0.999, 1e-15))
    extras.append({"kernel": f"k_adam (fields group, {n_par / 1e6:.1f} M parameters)", "bound": "hbm", "launch_us": us, "achieved": 28 * n_par / us / 1e3, "peak": peak,
                   "unit": "GB/s", "frac": 28 * n_par / us / 1e3 / peak, "algorithmic_bytes_per_param": 28})
    del p2, m2, v2, g2
    extras.extend(field_mlp_roofline(torch, nv, model.field, N, dev, timed_us, tpeak, tpeak_src, x=x, S=model.config.num_nerf_samples_per_ray))
    roofline["others"] = extras
    return roofline


def field_mlp_roofline(torch, nv, fld, N, dev, timed_us, tpeak, tpeak_src, x=None, S=48):
    """The field networks' tensor-core kernels on N samples against the dense fp16/bf16 tensor peak (SURVEY 8d FLOP counts, unpadded).
    First the two fused launches the mapping step runs (csrc/field_tc.cu: mlp_base + normals chain + assembly + mlp_head + mlp_pred_normals
    forward; mlp_head + assembly + mlp_base backward with dgrad and wgrad), then the per-network kernels behind the tcnn-style modules."""
    out = []
    ops = nv.ops
    if x is not None and N % S == 0:
        B = N // S
        enc = fld.mlp_base.encoder
        flat = lambda ps: ops._flat_of([p.detach() for p in ps])
        with torch.no_grad():
            feat16, jac = ops.grid_forward_jac(x, enc.hash_table.detach(), enc.spec)
            pn_params = fld.mlp_pred_normals._flat_param_list() + [fld.field_head_pred_normals.net.weight, fld.field_head_pred_normals.net.bias]
            img = ops.field_pack_weights(enc.spec, flat(fld.mlp_base.mlp._flat_param_list()), flat(fld.mlp_head._flat_param_list()), flat(pn_params))
            g = torch.Generator(device=dev).manual_seed(5)
            dirs = torch.nn.functional.normalize(torch.randn(B, 3, device=dev, generator=g), dim=-1)
            pos = (x * 4 - 2).contiguous()
            cam = torch.zeros(B, dtype=torch.int64, device=dev)
            emb = fld.embedding_appearance.embedding.weight.detach()
            sel = torch.ones(N, device=dev)
            fwd = lambda: ops.field_forward(feat16, jac, pos, dirs, cam, emb, sel, img, B, S, True, True)
            density, rgb, pn, normals, h0, pn_raw, saved = fwd()
            us_f = timed_us(fwd)
            dbase = torch.zeros(sum(p.numel() for p in fld.mlp_base.mlp._flat_param_list()), device=dev)
            dhead = torch.zeros(sum(p.numel() for p in fld.mlp_head._flat_param_list()), device=dev)
            demb = torch.zeros_like(emb)
            dden = torch.randn(N, device=dev, generator=g) * 1e-3
            drgb = torch.randn(N, 3, device=dev, generator=g) * 1e-3
            us_b = timed_us(lambda: ops.field_backward(feat16, saved, False, img, rgb, h0, sel, cam, dden, drgb, None, B, S, dbase, dhead, demb))
        # forward: the three networks (43 008 FLOP) + the normals chain's input-gradient product (2 x 64 x 32); backward: dgrad + wgrad of
        # mlp_head and mlp_base = 2 x (16 640 + 6 144) (mlp_pred_normals receives no gradient in NeRF-VO: pred_normal_loss_mult = 0)
        for tag, t_us, fl, note in (("k_field_fwd (mlp_base + normals chain + assembly + mlp_head + mlp_pred_normals, activations saved)", us_f, 43008 + 4096,
                                     "HBM per sample: 64 B features + 192 B feature derivatives read, 512 B activations + 44 B outputs written"),
                                    ("k_field_bwd (mlp_head + assembly + mlp_base: dgrad + wgrad)", us_b, 2 * (16640 + 6144),
                                     "HBM per sample: 512 B saved activations + 64 B features read, 128 B fp32 feature gradient written")):
            tf = fl * N / t_us / 1e6
            out.append({"kernel": tag, "bound": "tensor", "launch_us": t_us, "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                        "flop_per_sample": fl, "peak_source": tpeak_src, "note": note})
        del feat16, jac, saved
    nets = {"mlp_base 32-64-16": (fld.mlp_base.mlp.spec, fld.mlp_base.mlp._flat_param_list(), 6144),
            "mlp_head 63-64-64-3": (fld.mlp_head.spec, fld.mlp_head._flat_param_list(), 16640)}
    for name, (spec, params, flop) in nets.items():
        with torch.no_grad():
            xin = torch.randn(N, spec.in_dim, device=dev)
            x16 = nv.ops.cast_pad_f16(xin, spec)
            wimg = nv.ops.tc_pack_weights(nv.ops._flat_of(params), spec)
            y, saved = nv.ops.mlp_tc_forward(x16, wimg, spec, N, True)
            dy = torch.randn_like(y)
            dflat = torch.zeros(spec.n_params, dtype=torch.float32, device=dev)
            us_f = timed_us(lambda: nv.ops.mlp_tc_forward(x16, wimg, spec, N, True))
            us_b = timed_us(lambda: nv.ops.mlp_tc_backward(x16, wimg, saved, y, dy, spec, True, True, dflat, dy_absmax=8.0))
        for tag, t_us, fl in (("fwd", us_f, flop), ("bwd (dgrad + wgrad)", us_b, 2 * flop)):
            tf = fl * N / t_us / 1e6
            out.append({"kernel": f"k_mlp_tc_{tag}: {name} (per-network kernel behind tcnn_api.Network; not in the fused step)", "bound": "tensor", "launch_us": t_us,
                        "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "flop_per_sample": fl, "peak_source": tpeak_src})
    return out


# ---------------------------------------------------------------------------------------------------------------------
# config 5: evaluation frames
# ---------------------------------------------------------------------------------------------------------------------
def run_eval_frame(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist

    import nerf_vo_b200 as nv

    rank, world, local, dev = _dist_setup()
    nv._lib.load()
    torch.manual_seed(0)
    m = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(), num_train_data=cfg["images"]).to(dev).eval()
    with torch.no_grad():
        m.field.mlp_base.encoder.hash_table.normal_(0, 0.1)  # 'trained-like' table: densities are not ~1 everywhere
        if world > 1:
            for p in m.parameters():
                dist.broadcast(p.data, 0)
    H, W = cfg["frame"]
    r = nv.NerfstudioRenderer(model=m, num_rays_per_chunk=args.chunk)
    intr = {"fx": 600.0, "fy": 600.0, "cx": 599.5, "cy": 339.5, "height": H, "width": W}

    def ext(i):
        e = np.eye(4)
        e[:3, 3] = [0.05 * (i % 16), 0.0, 0.02 * (i % 16)]
        return e

    rows = None if world == 1 else nv.sharding.row_shard(H, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for s in range(steps):
            fn(s)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    n0 = nv._lib.launch_count()
    for s in range(max(3, args.warmup)):
        r.render_frame_device(intr, ext(s), rows=rows)
    per_frame = (nv._lib.launch_count() - n0) // max(3, args.warmup)
    with ClockSampler(local) as clk:
        ms = timed(lambda s: r.render_frame_device(intr, ext(s), rows=rows), args.steps)  # this rank's rows, result left on the device
    out = {}

    def full(s):
        out["c"], out["d"] = r.render_frame(intr, ext(s))  # pose in from the host, the finished frame (uint8 colour + fp32 depth) back on the host

    for s in range(2):
        full(s)
    ms_e2e = timed(full, args.steps)
    rays = H * W
    roofline = None
    if rank == 0 and not args.no_roofline:
        peak, peak_src = load_peaks()
        enc = m.field.mlp_base.encoder
        n_s = 65536 * 48
        x = torch.rand(n_s, 3, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        us = _timed_us(torch, flush, lambda: nv.ops.grid_forward(x, enc.hash_table.detach(), enc.spec, "tmh"))
        b_s = 12 + 16 * 8 * 2 * 4 + 16 * 2 * 2
        roofline = {"bound": "hbm", "kernel": "main hash grid forward (fp32 table -> fp16 TMH tiles), one 65536-ray chunk = 3.1 M samples, uniform positions",
                    "achieved": b_s * n_s / us / 1e3, "peak": peak, "unit": "GB/s", "frac": b_s * n_s / us / 1e3 / peak, "traffic": None, "peak_source": peak_src,
                    "launch_us": us, "algorithmic_bytes_per_launch": b_s * n_s}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, sec = cpu_port_frame_rays_per_s(16384, 2, 1, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "cpu_model": cpu_model_name(), "kind": "port",
               "sample": "2 timed passes (1 warm-up) x 16384 rays of the frame, eval-mode forward in 4096-ray chunks, oracle port of the reference torch path"}
    if rank == 0:
        line = {
            "metric": "nerf_eval_render_rays_per_s", "value": rays * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": f"{cfg['name']}: {rays} rays per frame, {world} GPU(s), image rows sharded, {args.chunk}-ray chunks, eval mode",
                       "baseline_config": 5, "rows_per_gpu": H if rows is None else rows[1] - rows[0], "l2": "64 MiB main table + per-chunk activations exceed nothing: the frame's "
                       "3.1 M samples per chunk stream 1.2 GB of activations, larger than the 126 MB L2"},
            "e2e": {"value": rays * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 16 * 4 + 4 * 4, "d2h_bytes_per_step": rays * (3 + 4),
                    "ms_per_step": ms_e2e / args.steps, "path": "NerfstudioRenderer.render_frame(intrinsics, extrinsics) -> (uint8 colour, float32 depth) numpy arrays"
                    + ("; rows all-gathered over NCCL" if world > 1 else "")},
            "gpu_launches": int(per_frame * args.steps), "gpu_launches_per_step": int(per_frame),
            "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "frame_checksum": int(out["c"].astype(np.int64).sum()),
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE workload (1-based as in SURVEY section 8): 2 = Replica step (default), "
                    "3 = 65536-ray ScanNet step, 4 = 2^21 table / 262144 global rays, 5 = evaluation frame")
    ap.add_argument("--mode", default=None, choices=["train", "eval-frame"], help="eval-frame = --config 5")
    ap.add_argument("--rays", type=int, default=0, help="override: rays per GPU per step")
    ap.add_argument("--log2-hashmap", type=int, default=0, help="override: log2 of the main table's rows per level")
    ap.add_argument("--chunk", type=int, default=1 << 16, help="eval-frame: rays per chunk (the reference uses 4096)")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pose-opt", action="store_true", help="headline step with the model's camera optimizer on (default: reported as the `pose_opt` leg)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the isolated-kernel roofline legs")
    ap.add_argument("--no-schedule-leg", action="store_true", help="skip the extra leg that follows the reference's proposal update schedule")
    ap.add_argument("--defer-fields", default="auto", choices=["auto", "on", "off"],
                    help="MappingTrainer(defer_fields_update=...): the fields group's optimizer / exchange of step k runs next to step k+1's proposal "
                         "sampling (auto = on for N > 1, where it hides the NVLink exchange)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N>1 gradient exchange: 'fused' = peer-memory reduce-scatter+Adam+all-gather kernel, 'nccl' = all-reduce + replicated Adam")
    args = ap.parse_args()
    claim_stdout()
    if args.mode == "eval-frame":
        args.config = 5
    cfg = PRESETS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.config == 5:
        run_eval_frame(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
