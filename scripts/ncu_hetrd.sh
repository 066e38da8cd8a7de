#!/bin/bash
# ncu capture of one tridiagonalisation panel kernel (symmetric path forced), n = 4096 real
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/eig_one.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
from neo_ls_svm_b200 import _lib
n = 4096
rng = np.random.default_rng(0)
Xt = rng.standard_normal((n, 48)) * 0.35
y = rng.standard_normal(n); s = np.full(n, 1.0 / n); sn = s / np.median(s)
gam = np.logspace(-6, np.log10(20), 128)
up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ctx = _lib.Context(0); ctx.set_eigensolver("dc")
ctx.dual_sweep(up(Xt), up(y), up(s), up(sn), up(gam), False)
torch.cuda.synchronize()
PY
NLS_HETRD_SYM_MIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hetrd_panel --launch-skip 30 --launch-count 1 -o gpurun_out/r2_ncu_hetrd_sym -f python /tmp/eig_one.py > gpurun_out/r2_ncu_hetrd_sym.log 2>&1
NLS_HETRD_SYM_MIN=100000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:hetrd_panel --launch-skip 30 --launch-count 1 -o gpurun_out/r2_ncu_hetrd_full -f python /tmp/eig_one.py > gpurun_out/r2_ncu_hetrd_full.log 2>&1
tail -3 gpurun_out/r2_ncu_hetrd_sym.log; ls -la gpurun_out/*.ncu-rep
