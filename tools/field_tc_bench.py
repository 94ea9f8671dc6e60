"""Fused field kernels (csrc/field_tc.cu) timed alone against the per-network launches they replace: CUDA events, L2 flushed between launches.
usage: python tools/field_tc_bench.py [rays ...]   (48 samples per ray)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nerf_vo_b200 as nv
import test_field_tc as T
from nerf_vo_b200.fields import FieldHeadNames as F

DEV = "cuda:0"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1590.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)

def timed_us(fn, reps=10):
    ev = []
    for i in range(3 + reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev[3:]) / reps * 1e3

out = []
for B in [int(a) for a in sys.argv[1:]] or [4096, 65536]:
    S = 48
    field = T._field(nv, K=64, log2=19).to(DEV).train()
    _, _, rs = T._samples(nv, field, B, S, K=64)
    ops = nv.ops
    fr = rs.frustums
    positions = fr.get_positions().reshape(-1, 3).contiguous()
    x, sel = ops.contract_normalize(positions)
    enc = field.mlp_base.encoder
    feat16, jac = ops.grid_forward_jac(x, enc.hash_table.detach(), enc.spec)
    flat = lambda ps: ops._flat_of([p.detach() for p in ps])
    pn_params = field.mlp_pred_normals._flat_param_list() + [field.field_head_pred_normals.net.weight, field.field_head_pred_normals.net.bias]
    img = ops.field_pack_weights(enc.spec, flat(field.mlp_base.mlp._flat_param_list()), flat(field.mlp_head._flat_param_list()), flat(pn_params))
    dirs = fr.directions.reshape(B, 3).contiguous()
    cam, emb = rs.camera_indices.reshape(B).long().contiguous(), field.embedding_appearance.embedding.weight.detach()
    n = B * S
    row = {"rays": B, "samples": n}
    for save in (True, False):
        us = timed_us(lambda: ops.field_forward(feat16, jac, positions, dirs, cam, emb, sel, img, B, S, True, save))
        flop = 43008 + 2 * 64 * 32  # SURVEY 8d forward FLOP of the three networks + the normals chain's dgrad product (64 x 32)
        row[f"fused_fwd_us{'_saving' if save else ''}"] = us
        row[f"fused_fwd_tflops{'_saving' if save else ''}"] = flop * n / us / 1e6
        row[f"fused_fwd_frac_of_tensor_peak{'_saving' if save else ''}"] = flop * n / us / 1e6 / peak
    with torch.no_grad():
        us_old = timed_us(lambda: field.forward(rs, compute_normals=True), reps=5)
    row["per_network_forward_us (contract + grid + 5 MLP launches + assembly + jac_dx, eager)"] = us_old
    us_grid = timed_us(lambda: ops.grid_forward_jac(x, enc.hash_table.detach(), enc.spec))
    row["grid_forward_jac_us"] = us_grid
    out.append(row)
    print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "field_tc_bench.json"), "w"), indent=1)
