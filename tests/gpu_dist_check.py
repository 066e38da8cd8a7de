"""Multi-GPU parity check, run under torchrun (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_dist_check.py

Rows of the golden case are sharded contiguously over the ranks; the sharded solve must select the
reference's γ index and reproduce β̂ / the LOO vectors to 1e-9, and agree with the 1-GPU solve to 1e-12.
"""

import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neo_ls_svm_b200 import _lib, _primal  # noqa: E402
from neo_ls_svm_b200.datasets import load_case  # noqa: E402


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = _lib.context(local)
    ok = True
    for name in ("c1", "clf_small", "c3_small"):
        with np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")) as z:
            g = {k: z[k] for k in z.files}
        X, y, sw, _, _ = load_case(name)
        classifier = bool(g["classifier"])
        y_ = np.where(y == np.unique(y)[0], -1.0, 1.0) if classifier else y.astype(np.float64)
        s = np.ones(len(y)) if sw is None else sw.astype(np.float64)
        s = s / s.sum()
        n = len(y)
        r0, r1 = rank * n // world, (rank + 1) * n // world
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)  # noqa: E731
        shift, W = up(g["shift"].ravel()), up(g["A_map"] / g["scale"].reshape(-1, 1))
        ctx.set_chunk_rows(1024)
        fit = _primal.primal_fit(up(X[r0:r1]), up(y_[r0:r1]), up(s[r0:r1]), shift, W, classifier, n_global=n, ctx=ctx)
        parts = [torch.zeros(n * (r + 1) // world - n * r // world, dtype=torch.float64, device=dev) for r in range(world)]
        gathered = {}
        for key in ("loo_residuals", "loo_std", "residuals", "loo_leverage"):
            dist.all_gather(parts, fit.rows[key].contiguous())
            gathered[key] = torch.cat(parts).cpu().numpy()
        if rank == 0:
            checks = {
                "opt": fit.opt == int(g["opt"]),
                "beta": rel(fit.beta.cpu().numpy(), g["beta"]) < 1e-9,
                "loo_errors": rel(fit.loo_errors, g["loo_errors"]) < 1e-9,
                "loo_score": abs(fit.loo_score - float(g["loo_score"])) < 1e-9,
            }
            for key in gathered:
                checks[key] = rel(gathered[key], g[key]) < 1e-9
            print(f"[dist] world={world} {name}: " + " ".join(f"{k}={'ok' if v else 'FAIL'}" for k, v in checks.items()), flush=True)
            ok = ok and all(checks.values())
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
