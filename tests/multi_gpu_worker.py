"""Worker of tests/test_multi_gpu.py: run under torchrun with N >= 2 ranks, one per GPU (NCCL).

Every rank builds the same model, takes ITS slice of each global batch (split_batches=True semantics: utils/common.py:61-65,
reconstruction.py:134-137), runs the CUDA kernels and the ONE NCCL all-reduce of the flat gradient arena, and checks the averaged
gradients against `ddp_emulated_grads` (float64 oracle: per-rank losses on the sub-batches, gradients averaged) -- for the eager
step, for the CUDA-graph step (the collective is captured inside the graph) and, over several Adam steps, that all ranks hold
bit-identical parameters.  Exits non-zero on any mismatch.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import rel                                               # noqa: E402
from oracle.ptycho_torch import ddp_emulated_grads                    # noqa: E402  (checker only)
from ptyrad_b200 import PtychoAD, CombinedLoss                        # noqa: E402
from ptyrad_b200.optim import FusedAdam                               # noqa: E402
from ptyrad_b200.step import GradArena, GraphedStep, recon_batch, shard_indices   # noqa: E402
from workloads import CONFIGS, make_inputs                            # noqa: E402

TOL = {"obja": 1e-4, "objp": 1e-4, "probe": 1e-4, "probe_pos_shifts": 3e-4, "obj_tilts": 5e-4}


class NoStep(FusedAdam):
    def step(self, closure=None):
        pass


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fails = []
    for cfg_name, nglobal in (("T64", 10), ("C2d", 8), ("T256", 6)):
        cfg = CONFIGS[cfg_name]
        iv, mp, lp = make_inputs(cfg, seed=71)
        rng = np.random.default_rng(2)
        batch = np.sort(rng.choice(cfg.scan ** 2, nglobal, replace=False)).astype(np.int64)
        mine = shard_indices(batch, rank, world)
        ref = ddp_emulated_grads(iv, mp, lp, batch, world, torch.float64)
        # Shift / tilt sums of a handful of positions nearly cancel; where float32 itself cannot resolve them to TOL (T256, 3 positions per
        # rank: the float32 run of the same expressions is 5.2e-4 away from float64) the bound is twice the float32 restatement's error.
        r32 = ddp_emulated_grads(iv, mp, lp, batch, world, torch.float32)
        tol = {k: (max(TOL[k], 2.0 * rel(np.asarray(r32[k], np.float64), g)) if k in ("probe_pos_shifts", "obj_tilts") else TOL[k])
               for k, g in ref.items()}
        for mode in ("eager", "graph"):
            model = PtychoAD(iv, mp, device=dev, verbose=False)
            loss_fn = CombinedLoss(lp, device=dev)
            arena = GradArena(model)
            opt = NoStep(model.optimizable_params)
            if mode == "eager":
                recon_batch(model, loss_fn, opt, mine, arena, world)
            else:
                GraphedStep(model, loss_fn, opt, arena, len(mine), world=world)(mine)
            torch.cuda.synchronize()
            for k, g in ref.items():
                e = rel(model.optimizable_tensors[k].grad.cpu().numpy(), g)
                ok = e < tol[k]
                if rank == 0:
                    print(f"{cfg_name:5s} {mode:5s} world={world} grad {k:18s} {e:.2e} (bound {tol[k]:.1e}) {'ok' if ok else 'FAIL'}", flush=True)
                if not ok:
                    fails.append((cfg_name, mode, k, e))
        # several real Adam steps through the graph: all ranks must end with bit-identical parameters
        model = PtychoAD(iv, mp, device=dev, verbose=False)
        loss_fn = CombinedLoss(lp, device=dev)
        arena = GradArena(model)
        opt = FusedAdam(model.optimizable_params)
        step = GraphedStep(model, loss_fn, opt, arena, len(mine), world=world)
        for s in range(4):
            b = np.sort(np.random.default_rng(10 + s).choice(cfg.scan ** 2, nglobal, replace=False)).astype(np.int64)
            step(shard_indices(b, rank, world))
        torch.cuda.synchronize()
        for k, t in model.optimizable_tensors.items():
            if not t.requires_grad:
                continue
            lo, hi = t.detach().clone(), t.detach().clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            if not torch.equal(lo, hi):
                fails.append((cfg_name, "params differ across ranks", k, float((hi - lo).abs().max())))
    torch.cuda.synchronize()
    dist.barrier()
    if fails:
        print(f"rank {rank}: FAILURES {fails}", flush=True)
    # captured NCCL graphs can dead-lock destroy_process_group(): leave through os._exit after the barrier
    sys.stdout.flush()
    os._exit(1 if fails else 0)


if __name__ == "__main__":
    main()
