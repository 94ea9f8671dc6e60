#!/usr/bin/env python
"""Kernel timeline of the mapping step (CUDA-graph replay) through torch.profiler/CUPTI: per kernel stream, start, duration.
Writes gpurun_out/timeline_<tag>.csv (one replay) and prints the per-stream busy time and the gaps on the capture stream.
usage: python tools/timeline.py [--rays 4096] [--tag s6]"""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_vo_b200 as nv
from nerf_vo_b200.synthetic import synthetic_jitters, synthetic_rays
from nerf_vo_b200.trainer import MappingTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=4096)
ap.add_argument("--tag", default="step")
ap.add_argument("--pose", default="SO3xR3")
a = ap.parse_args()
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:  # under torchrun: every rank steps (fused exchange), rank 0 writes its own timeline
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
model = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(camera_optimizer_mode=a.pose), num_train_data=192).to(dev)
tr = MappingTrainer(model, num_rays=a.rays)
rays, targets = synthetic_rays(a.rays, num_images=192, seed=1234 + 1000 * rank)
jit = synthetic_jitters(a.rays, seed=99 + 1000 * rank)
tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in jit])
tr.capture(warmup=3)
for _ in range(5):
    tr.train_step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.train_step()
    torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    if rank != 0:
        dist.destroy_process_group()
        sys.exit(0)
path = os.path.join(ROOT, "gpurun_out", f"trace_{a.tag}.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
os.remove(path)
ev.sort(key=lambda e: e["ts"])
# split into replays by the largest gaps: take the LAST replay
n = len(ev) // 3
last = ev[-n:]
t0 = last[0]["ts"]
with open(os.path.join(ROOT, "gpurun_out", f"timeline_{a.tag}.csv"), "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in last:
        f.write(f"{e['ts'] - t0:.2f},{e['dur']:.2f},{e['args'].get('stream')},\"{e['name'][:60]}\"\n")
end = max(e["ts"] + e["dur"] for e in last) - t0
print(f"replay: {len(last)} activities, span {end:.1f} us")
