"""Single-GPU timing of the exchange kernel at world_size 1 (no peers) against nvo_adam_step on the same flat size:
separates the cost of system-scope loads/stores and the flag protocol from NVLink itself."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nerf_vo_b200 as nv
from nerf_vo_b200 import ops
from nerf_vo_b200.peer import PeerBuffers

dev = torch.device("cuda", 0)
n = 19_427_104
bufs = PeerBuffers(n, dev)
bufs.grads.normal_()
flat, g, m, v = (torch.randn(n, device=dev) for _ in range(4))
v.abs_()
sa, sb = torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("k_exchange_adam<1> us:", t(lambda: bufs.adam_exchange_step(sa, 1e-2, 0.9, 0.999, 1e-15)))
print("k_adam us:", t(lambda: ops.adam_step(flat, g, m, v, sb, 1e-2, 0.9, 0.999, 1e-15, 1.0)))
