import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_vo_b200 as nv
from nerf_vo_b200.synthetic import synthetic_rays, synthetic_jitters
from nerf_vo_b200.trainer import MappingTrainer

dev = torch.device("cuda:0")
def make():
    torch.manual_seed(0)
    cfg = nv.NerfactoModelConfig(log2_hashmap_size=14)
    model = nv.ExtendedNerfactoModel(cfg, num_train_data=16).to(dev)
    tr = MappingTrainer(model, num_rays=512, use_cuda_graph=False)
    rays, targets = synthetic_rays(512, num_images=16)
    tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in synthetic_jitters(512)])
    tr.capture(warmup=3)
    return tr

def attempt(name, fn):
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, "OK", flush=True)
    except Exception as e:
        print(name, "FAILED:", str(e).splitlines()[0], flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("sync err", e2)

tr = make()
tr.model.proposal_sampler._steps_since_update = 10**6
attempt("A: first capture = fwd_bwd (no eager step before)", tr._forward_backward)
tr = make()
tr.model.proposal_sampler._steps_since_update = 10**6
def both():
    tr._forward_backward(); tr._optimizer()
attempt("B: first capture = fwd_bwd+adam", both)
tr = make()
tr.train_step()
tr.model.proposal_sampler._steps_since_update = 10**6
attempt("C: eager step, then fwd_bwd+adam", both)
