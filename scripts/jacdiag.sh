for mode in both pivot update; do
echo "== $mode"
if [ $mode = both ]; then unset NLS_JACOBI_DIAG; else export NLS_JACOBI_DIAG=$mode; fi
python - <<'PY'
import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from neo_ls_svm_b200 import _lib
ctx=_lib.Context(0); ctx.set_eigensolver('jacobi')
rng=np.random.default_rng(0)
m=1025; n=4*m
Z=rng.standard_normal((n,m-1))*1.5
phi=np.concatenate([np.exp(-1j*Z)/np.sqrt(m-1), np.ones((n,1))],axis=1)/n
A=phi.conj().T@phi; A=(A+A.conj().T)/2
Ad=torch.from_numpy(A).cuda()
for rep in range(3):
    torch.cuda.synchronize(); t=time.perf_counter()
    try:
        ctx.heev(Ad, float(n)*m)
    except Exception as e: print('exc', str(e)[:80])
    torch.cuda.synchronize(); print('ms %.2f sweeps %d' % ((time.perf_counter()-t)*1e3, ctx.last_eig_sweeps()))
PY
done
