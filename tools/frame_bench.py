#!/usr/bin/env python
"""BASELINE config 5: evaluation full-frame render 1200x680 (816 000 rays, forward only) through NerfstudioRenderer.render_frame — the call
evaluation/nerf_renderer.py:132-168 makes per frame — timed end to end (host numpy arrays out), plus the device-only part.
usage: python tools/frame_bench.py [--frames 10] [--chunk 65536]      (under torchrun: rows sharded across ranks)"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import nerf_vo_b200 as nv

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=10)
ap.add_argument("--chunk", type=int, default=1 << 16)
a = ap.parse_args()
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = nv.ExtendedNerfactoModel(nv.NerfactoModelConfig(), num_train_data=192).to(dev).eval()
with torch.no_grad():
    m.field.mlp_base.encoder.hash_table.normal_(0, 0.1)
r = nv.NerfstudioRenderer(model=m, num_rays_per_chunk=a.chunk)
intr = {"fx": 600.0, "fy": 600.0, "cx": 599.5, "cy": 339.5, "height": 680, "width": 1200}
exts = []
for i in range(a.frames + 2):
    e = np.eye(4)
    e[:3, 3] = [0.05 * i, 0.0, 0.02 * i]
    exts.append(e)
for e in exts[:2]:
    r.render_frame(intr, e.copy())
torch.cuda.synchronize()
t0 = time.perf_counter()
for e in exts[2:]:
    color, depth = r.render_frame(intr, e.copy())
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / a.frames
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for e in exts[2:]:
    r.render_frame_device(intr, e.copy(), rows=None if world == 1 else nv.sharding.row_shard(680, rank, world))
ev1.record()
torch.cuda.synchronize()
dd = ev0.elapsed_time(ev1) / a.frames
if rank == 0:
    print(json.dumps({"workload": "evaluation frame 1200x680, forward only, eval mode", "n_gpus": world, "chunk_rays": a.chunk, "frames": a.frames,
                      "ms_per_frame_e2e_host_arrays": dt * 1e3, "rays_per_s_e2e": 816000 / dt, "ms_per_frame_device": dd, "rays_per_s_device": 816000 / (dd * 1e-3),
                      "color_shape": list(color.shape), "depth_dtype": str(depth.dtype)}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
