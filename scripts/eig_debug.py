"""Stage-by-stage check of the dc eigensolver on the GPU against the NumPy prototypes (scripts/proto)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts", "proto"))
import torch
from neo_ls_svm_b200 import _lib
from hetrd_proto import hetrd_blocked

def check(m, lo_exp, calls, seed=0):
    rng = np.random.default_rng(seed)
    U, _ = np.linalg.qr(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)))
    lamt = np.logspace(0, lo_exp, m)
    A = (U * lamt) @ U.conj().T; A = (A + A.conj().T) / 2
    ref = np.linalg.eigvalsh(A)
    d0, e0, *_ = hetrd_blocked(A)
    ctx = _lib.Context(0); ctx.set_eigensolver("dc")
    Ad = torch.from_numpy(A).cuda()
    for c in range(calls):
        lam, Q = ctx.heev(Ad, 1.0)
        lam, Q = lam.cpu().numpy(), Q.cpu().numpy()
        d, e = ctx.last_tridiagonal(m)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        tl = np.linalg.eigvalsh(T)
        lz, Zt = ctx.stedc(d, e)
        lz, Zt = lz.cpu().numpy(), Zt.cpu().numpy()
        print(f"m={m} lo={lo_exp} call {c}: d vs proto {np.max(np.abs(d-d0)):.1e} e vs proto {np.max(np.abs(np.abs(e)-np.abs(e0))):.1e} | "
              f"tridiag eig vs A {np.max(np.abs(tl-ref)):.1e} | stedc lam {np.max(np.abs(lz-tl)):.1e} orth {np.max(np.abs(Zt@Zt.T-np.eye(m))):.1e} "
              f"resid {np.max(np.abs(T@Zt.T-Zt.T*lz)):.1e} | full lam {np.max(np.abs(lam-ref)):.1e} orth {np.max(np.abs(Q.conj().T@Q-np.eye(m))):.1e} "
              f"resid {np.max(np.abs(A@Q-Q*lam)):.1e}", flush=True)

for m, lo in ((257, -14), (513, -12), (513, -14), (513, -16), (1025, -14)):
    check(m, lo, 2)
