"""Isolated launches of the tensor-core MLP kernels at the mapping-step shape (196608 samples) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nerf_vo_b200 as nv
from nerf_vo_b200 import ops

dev = "cuda:0"
n = 4096 * 48
torch.manual_seed(0)
specs = {"base": ops.MlpSpec(32, (64, 16), acts=("relu", "none")),
         "head": ops.MlpSpec(63, (64, 64, 3), acts=("relu", "relu", "sigmoid")),
         "pn": ops.MlpSpec(27, (64, 64, 64, 3), acts=("relu", "relu", "none", "tanh"))}
which = sys.argv[1:] or list(specs)
for name in which:
    spec = specs[name]
    flat = (torch.randn(spec.n_params, device=dev) * 0.1)
    x16 = ops.cast_pad_f16(torch.randn(n, spec.in_dim, device=dev) * 0.5, spec)
    dflat = torch.zeros(spec.n_params, device=dev)
    flat = ops.tc_pack_weights(flat, spec)
    for it in range(3):
        y, saved = ops.mlp_tc_forward(x16, flat, spec, n, True)
        dy = torch.randn_like(y)
        dx, _ = ops.mlp_tc_backward(x16, flat, saved, y, dy, spec, True, True, dflat)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for it in range(10):
        y, saved = ops.mlp_tc_forward(x16, flat, spec, n, True)
    e[1].record()
    for it in range(10):
        dx, _ = ops.mlp_tc_backward(x16, flat, saved, y, dy, spec, True, True, dflat)
    e[2].record()
    torch.cuda.synchronize()
    print(f"{name}: fwd {e[0].elapsed_time(e[1]) * 100:.1f} us  bwd {e[1].elapsed_time(e[2]) * 100:.1f} us  (n={n})", flush=True)
