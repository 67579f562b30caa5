"""Multi-GPU check (run under torchrun on a B200 box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_gpu_check.py
Clips sharded over ranks (NCCL), all-gathered per-clip MSE / SR frames must equal the single-GPU
result bit for bit and be in clip order (model/pfnl.py:90,139-141)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pfnl_b200 import PFNL, dist as D  # noqa: E402
from pfnl_b200 import weights as WT  # noqa: E402


def main():
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n_total, h, w = 6, 16, 24
    g = torch.Generator().manual_seed(7)
    lr = torch.rand((n_total, 7, h, w, 3), generator=g)
    hr = torch.rand((n_total, 1, 4 * h, 4 * w, 3), generator=g)
    m = PFNL(weights=WT.xavier_init(), device=local, precision="fp16x3")
    s, e = D.shard_range(n_total, rank, world)
    mse, psnr = D.sharded_eval_mse(m, lr[s:e].to(dev), hr[s:e].to(dev), n_total)
    sr_local = m.forward(lr[s:e].to(dev))
    sr_all = D.all_gather_clips(sr_local, n_total)
    # single-GPU reference on every rank
    sr_ref = m.forward(lr.to(dev))
    mse_ref = m.engine.mse(sr_ref, hr.to(dev))[:, None]
    ok = torch.equal(sr_all, sr_ref) and torch.equal(mse, mse_ref)
    psnr_ref = 10.0 * torch.log10(1.0 / mse_ref.double())
    ok = ok and torch.allclose(psnr, psnr_ref)
    print(f"rank {rank}/{world}: shard [{s},{e}) sharded==single-GPU: {ok}; mse[:3]={mse[:3, 0].tolist()}", flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
