"""2-rank diagnostic (not a test): deferred vs undeferred fields update on the fused peer-memory arm, eager and CUDA graph.
usage: python tools/diag_defer2.py   (spawns two ranks on cuda:0 / cuda:1)"""
import os, socket, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(rank, world, port):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import nerfacto_oracle as O
    import nerf_vo_b200 as nv
    from nerf_vo_b200.trainer import MappingTrainer
    K, B, STEPS = 8, 256, 4
    rays, targets = O.synthetic_rays(B, num_images=K, seed=3 + rank)
    jit = O.synthetic_jitters(B)
    for exchange in ("fused", "nccl"):
        for graph in (False, True):
            res = {}
            for defer in (False, True):
                if exchange == "nccl" and defer:
                    continue
                torch.manual_seed(0)
                cfg = nv.NerfactoModelConfig(log2_hashmap_size=14)
                for a in cfg.proposal_net_args_list:
                    a["log2_hashmap_size"] = 12
                model = nv.ExtendedNerfactoModel(cfg, num_train_data=K).to(dev)
                tr = MappingTrainer(model, num_rays=B, lr=1e-2, eps=1e-15, use_cuda_graph=graph, exchange=exchange, defer_fields_update=defer)
                init = tr.flat.detach().clone()
                if graph:
                    tr.capture(warmup=2)
                tr.set_inputs({k: v.to(dev) for k, v in rays.items()}, {k: v.to(dev) for k, v in targets.items()}, [j.to(dev) for j in jit])
                losses, moves = [], []
                for s in range(STEPS):
                    losses.append(float(tr.train_step()))
                    if s == 0:
                        tr.flush()
                        torch.cuda.synchronize()
                        dist.barrier()
                        moves.append((tr.flat.detach() - init).double().cpu())
                tr.flush()
                torch.cuda.synchronize()
                dist.barrier()
                moves.append((tr.flat.detach() - init).double().cpu())
                res[defer] = (losses, moves)
                dist.barrier()
            if rank == 0:
                l0 = res[False][0]
                print(f"{exchange} graph={graph} undeferred losses {['%.6f' % x for x in l0]}", flush=True)
                if True in res:
                    l1 = res[True][0]
                    print(f"{exchange} graph={graph}   deferred losses {['%.6f' % x for x in l1]}", flush=True)
                    for i, name in enumerate(("after step 1", f"after {STEPS} steps")):
                        a, b = res[False][1][i], res[True][1][i]
                        cos = float((a * b).sum() / (a.norm() * b.norm()))
                        print(f"    movement {name}: cos {cos:.6f} |a| {float(a.norm()):.5f} |b| {float(b.norm()):.5f} differing entries {int((a != b).sum())} of {a.numel()} max|a-b| {float((a-b).abs().max()):.3e}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(worker, args=(2, port), nprocs=2, join=True)
