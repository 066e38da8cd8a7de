"""A/B of the projection -> sweep plane fusion at the C5 shape (m = 4097) and at C3's (m = 1025): profiled kernel kinds."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neo_ls_svm_b200 import _lib, _primal
rng = np.random.default_rng(0)
for (n, d, D) in ((24000, 128, 4096), (65536, 64, 1024)):
    X = rng.standard_normal((n, d)); y = np.sin(X[:, 0]) + 0.1 * rng.standard_normal(n); s = np.full(n, 1.0 / n)
    W = rng.standard_normal((d, D)) * 0.3; shift = np.zeros(d)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    Xd, yd, sd, shd, Wd = dev(X), dev(y), dev(s), dev(shift), dev(W)
    for fuse in ("1", "0", "1", "0"):
        os.environ["NLS_OZ_FUSE"] = fuse
        ctx = _lib.context(0)
        ctx.profile(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        fit = _primal.primal_fit(Xd, yd, sd, shd, Wd, False, ctx=ctx)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        prof = ctx.profile_read(); ctx.profile(False)
        print(f"m={D+1} n={n} fuse={fuse}: solve {t*1e3:.1f} ms opt={fit.opt}", {k: round(v['ms'], 1) for k, v in prof.items()}, flush=True)
