import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nerf_vo_b200 as nv
from nerf_vo_b200.synthetic import synthetic_rays, synthetic_jitters
from nerf_vo_b200.trainer import MappingTrainer

dev = torch.device("cuda:0")
for (log2, B, K) in ((14, 512, 16), (19, 512, 192), (14, 4096, 192), (19, 4096, 192)):
    torch.manual_seed(0)
    cfg = nv.NerfactoModelConfig(log2_hashmap_size=log2)
    model = nv.ExtendedNerfactoModel(cfg, num_train_data=K).to(dev)
    tr = MappingTrainer(model, num_rays=B, use_cuda_graph=True)
    rays, targets = synthetic_rays(B, num_images=K)
    tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in synthetic_jitters(B)])
    try:
        tr.capture(warmup=3)
        l = float(tr.train_step())
        print(log2, B, K, "capture OK loss", l, "launches", tr.launches_per_step)
    except Exception as e:
        print(log2, B, K, "FAILED", str(e).splitlines()[0])
        break
    del tr, model
    torch.cuda.empty_cache()
