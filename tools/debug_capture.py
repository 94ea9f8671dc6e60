import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_vo_b200 as nv
from nerf_vo_b200.synthetic import synthetic_rays, synthetic_jitters
from nerf_vo_b200.trainer import MappingTrainer

dev = torch.device("cuda:0")
torch.manual_seed(0)
cfg = nv.NerfactoModelConfig(log2_hashmap_size=14)
model = nv.ExtendedNerfactoModel(cfg, num_train_data=16).to(dev)
B = 512
tr = MappingTrainer(model, num_rays=B, use_cuda_graph=False)
rays, targets = synthetic_rays(B, num_images=16)
tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in synthetic_jitters(B)])
tr.capture(warmup=3)
print("eager ok, loss", float(tr.train_step()))

def attempt(name, fn, mode):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, mode, "OK")
    except Exception as e:
        print(name, mode, "FAILED:", str(e).splitlines()[0])
        torch.cuda.synchronize()

m = tr.model
def fwd_only():
    with torch.no_grad():
        i = tr.inputs
        m.proposal_sampler._steps_since_update = 10**6
        m(tr._bundle(), [i["jitter0"], i["jitter1"], i["jitter2"]])
def fwd_grad():
    i = tr.inputs
    m.proposal_sampler._steps_since_update = 10**6
    batch = {"image": i["rgb"], "depth_image": i["depth"], "normal_image": i["normal"]}
    m.get_train_loss_dict(tr._bundle(), batch, [i["jitter0"], i["jitter1"], i["jitter2"]])
for mode in ("global", "relaxed"):
    attempt("zero", lambda: tr.grad.zero_(), mode)
    attempt("adam", tr._optimizer, mode)
    attempt("fwd_nograd", fwd_only, mode)
    attempt("fwd_grad", fwd_grad, mode)
    attempt("fwd_bwd", tr._forward_backward, mode)
