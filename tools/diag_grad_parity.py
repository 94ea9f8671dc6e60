"""Diagnostic: where does the fp32 path's table / base-layer-0 gradient leave the oracle at config-2 size?  GPU vs oracle fp32 vs oracle fp64."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import nerfacto_oracle as O
import nerf_vo_b200 as nv
import test_full_size_parity as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
ocfg, P, m, rays, targets, jit, rb, batch = T._setup(nv, 19, 17, 192, prec, B)
rec = {}
orig_gb = nv.ops.grid_backward
def gb(x, dy, spec, dtable=None, n_rows=None, tmf=False):
    if spec.n_levels == 16:
        rec["x"], rec["dy"], rec["tmf"] = x.detach().clone(), dy.detach().clone(), tmf
    return orig_gb(x, dy, spec, dtable=dtable, n_rows=n_rows, tmf=tmf)
nv.ops.grid_backward = gb
outputs, ld, _ = m.get_train_loss_dict(rb, batch, [j.to("cuda") for j in jit])
sum(ld.values()).backward()
torch.cuda.synchronize()
torch.set_num_threads(os.cpu_count())

def run_oracle(dtype):
    store = []
    orig = O.hash_encode
    def he(x, table, sc, log2):
        y = orig(x, table, sc, log2)
        if y.requires_grad:
            y.retain_grad(); store.append(y)
        return y
    O.hash_encode = he
    Pg = {k: v.clone().to(dtype).requires_grad_(True) for k, v in P.items()}
    r = {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in rays.items()}
    t = {k: v.to(dtype) for k, v in targets.items()}
    if dtype == torch.float64:
        torch.set_default_dtype(torch.float64)
    try:
        oout, oL, tot = O.mapping_step(Pg, ocfg, r, t, [j.to(dtype) for j in jit])
    finally:
        torch.set_default_dtype(torch.float32)
        O.hash_encode = orig
    main = [s for s in store if s.shape[1] == 32][-1]
    return Pg, main.grad.detach(), oout

P32, df32, o32 = run_oracle(torch.float32)
P64, df64, o64 = run_oracle(torch.float64)
dy = rec["dy"]
n = B * 48
if rec["tmf"]:
    tiles = (n + 127) // 128
    dy = dy.view(tiles, 32, 128).permute(0, 2, 1).reshape(-1, 32)[:n]
dy = dy.cpu()
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())
out = {"dfeat gpu vs o32": rel(dy, df32), "dfeat gpu vs o64": rel(dy, df64), "dfeat o32 vs o64": rel(df32, df64)}
out["x max abs diff gpu-o32"] = float((rec["x"].cpu() - o32["field"]["x_normalized"].reshape(-1, 3)).abs().max())
out["x max abs diff o32-o64"] = float((o64["field"]["x_normalized"].reshape(-1, 3) - o32["field"]["x_normalized"].reshape(-1, 3).double()).abs().max())
for name in ("field.mlp_base.model.0.hash_table", "field.mlp_base.model.1.layers.0.weight", "field.mlp_base.model.1.layers.1.weight", "field.mlp_head.layers.0.weight",
             "proposal_networks.1.encoding.hash_table"):
    g = dict(m.named_parameters())[name].grad.cpu()
    out[name] = {"gpu vs o32": rel(g, P32[name].grad), "gpu vs o64": rel(g, P64[name].grad), "o32 vs o64": rel(P32[name].grad, P64[name].grad)}
g = dict(m.named_parameters())["field.mlp_base.model.0.hash_table"].grad.cpu().view(16, -1, 2)
g32, g64 = P32["field.mlp_base.model.0.hash_table"].grad.view(16, -1, 2), P64["field.mlp_base.model.0.hash_table"].grad.view(16, -1, 2)
mx = float(g64.abs().max())
out["table per level (err / global max)"] = [{"lvl": l, "gpu-o64": float((g[l].double() - g64[l]).abs().max()) / mx, "o32-o64": float((g32[l].double() - g64[l]).abs().max()) / mx,
                                               "level max / global max": float(g64[l].abs().max()) / mx} for l in range(16)]
dt = orig_gb(o32["field"]["x_normalized"].reshape(-1, 3).float().cuda().contiguous(), df64.float().cuda().contiguous(), m.field.mlp_base.encoder.spec).cpu()
out["scatter alone (oracle32 x, oracle dfeat64) vs o64"] = rel(dt.view(-1, 2), P64["field.mlp_base.model.0.hash_table"].grad)
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"diag_grad_parity_{prec}_{B}.json"), "w"), indent=1)
