"""Accuracy of the parallel-in-time Cholesky / UDU / naturals paths against the sequential sweeps on
ill-conditioned (small dt / lengthscale) Matern posterior precisions, measured against a long-double
factorisation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from markovflow_b200 import SymmetricBlockTriDiagonal as S, _lib
from oracle import np_oracle as O

lib = _lib.lib()
dev = torch.device("cuda:0")


def chol_longdouble(diag, sub):
    """block Cholesky in np.longdouble (x87 80-bit): reference for the error of both float64 paths."""
    diag, sub = diag.astype(np.longdouble), sub.astype(np.longdouble)
    t, d = diag.shape[0], diag.shape[-1]
    ld, ls = np.zeros_like(diag), np.zeros_like(sub)
    prev = None
    for k in range(t):
        s = diag[k].copy()
        if prev is not None:
            s -= prev @ prev.T
        l = np.zeros((d, d), dtype=np.longdouble)
        for j in range(d):
            l[j, j] = np.sqrt(s[j, j] - (l[j, :j] ** 2).sum())
            for i in range(j + 1, d):
                l[i, j] = (s[i, j] - (l[i, :j] * l[j, :j]).sum()) / l[j, j]
        ld[k] = l
        if k + 1 < t:
            # Ls = A L^-T
            a = sub[k]
            x = np.zeros((d, d), dtype=np.longdouble)
            for j in range(d):
                x[:, j] = (a[:, j] - x[:, :j] @ l[j, :j]) / l[j, j]
            ls[k] = x
            prev = x
    return ld, ls


for name, kern, ratio in (("Matern32", O.Matern32, 0.2), ("Matern32", O.Matern32, 0.01), ("Matern52", O.Matern52, 0.2),
                          ("Matern52", O.Matern52, 0.02), ("Matern52", O.Matern52, 0.005)):
    rng = np.random.default_rng(3)
    t = 4000
    ell = 1.0
    tp = np.cumsum(ell * ratio * rng.uniform(0.5, 1.5, size=t))
    k = kern(ell, 1.0)
    diag, sub = O.kalman_k_inv_post(k.state_space_model(tp), k.emission_matrix(tp), np.array([[100.0]]))
    ref_ld, ref_ls = chol_longdouble(diag, sub)
    out = {}
    for knob in (0, 1):
        lib.mf_set_tuning(2, knob)
        c = S(torch.as_tensor(diag[None], device=dev), torch.as_tensor(sub[None], device=dev)).cholesky
        out[knob] = (c.block_diagonal[0].cpu().numpy(), c.block_sub_diagonal[0].cpu().numpy())
    lib.mf_set_tuning(2, 0)
    err = lambda a, r: float(np.max(np.abs(a - r)) / np.max(np.abs(r)))
    print(f"{name} dt/ell~{ratio}: cond(diag blocks) ~{np.linalg.cond(diag[t // 2]):.1e} | "
          f"PIT err Ld {err(out[0][0], ref_ld):.1e} Ls {err(out[0][1], ref_ls):.1e} | "
          f"sequential err Ld {err(out[1][0], ref_ld):.1e} Ls {err(out[1][1], ref_ls):.1e}", flush=True)
