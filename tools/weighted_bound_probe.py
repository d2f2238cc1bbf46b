#!/usr/bin/env python
"""How sharp is the weighted Samuelson bound of the default screen (rmsd_screen.cu, ScFrame) on synthetic ensembles of
several shapes?  numpy only: for random pairs, the bound sqrt(sum_b w_b ||col_b(S)||^2) in the principal-axes frame against
lambda_max and against the threshold eigenvalue, for three weight choices.  python tools/weighted_bound_probe.py"""
import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tscode_b200.synth import gen_ensemble
def study(name, scale, N=1500, M=80, rot=None, sigma_cluster=1.0):
    S = gen_ensemble(3, N, M, N//10, scale=np.array(scale) if not np.isscalar(scale) else scale, sigma_cluster=sigma_cluster)
    S = S - S.mean(1, keepdims=True)
    if rot is not None:
        S = S @ rot.T
    # principal frame from the first structure
    X0 = S[0]
    lam, Q = np.linalg.eigh(X0.T @ X0)
    Sp = S @ Q            # coordinates in principal frame (columns = eigvecs)
    c = np.sqrt(np.maximum(lam, 1e-12))      # expected sqrt f_a ratios ~ second moments? use lam (s_a ~ sum x_a^2)
    thr = 0.5
    G = (S**2).sum((1,2))
    rng = np.random.default_rng(0)
    I = rng.integers(0, N, 20000); J = rng.integers(0, N, 20000)
    keep = I != J; I, J = I[keep], J[keep]
    cov = np.einsum('nma,nmb->nab', Sp[I], Sp[J])
    sv = np.linalg.svd(cov, compute_uv=False)
    det = np.linalg.det(cov)
    lmax = sv[:,0]+sv[:,1]+np.where(det>=0, sv[:,2], -sv[:,2])
    lam_t = 0.5*(G[I]+G[J]) - 0.5*M*thr*thr
    similar = lmax >= lam_t
    fa = (cov**2).sum(2)        # row norms squared, rows = i-side component a
    for wname, cc in (("lam", lam), ("sqrtlam", np.sqrt(lam)), ("iso", np.ones(3))):
        w = cc.sum()/cc
        bound = np.sqrt((fa*w).sum(1))
        excl = bound < lam_t
        print(f"{name:10s} weights {wname:8s} w={np.round(w,2)} non-similar {int((~similar).sum())} excluded {int(excl.sum())} "
              f"({100*excl.sum()/max(1,(~similar).sum()):.2f} %) unsound {int((excl & similar).sum())}  median bound/lmax {np.median(bound/lmax):.4f} max-needed {np.median(lam_t/lmax):.4f}")
study("isotropic", 3.0)
study("elongated", [6.0,2.0,1.0])
study("planar", [4.0,4.0,0.5])
th=0.7; R=np.array([[np.cos(th),-np.sin(th),0],[np.sin(th),np.cos(th),0],[0,0,1.0]])@np.array([[1,0,0],[0,np.cos(0.4),-np.sin(0.4)],[0,np.sin(0.4),np.cos(0.4)]])
study("elong-rot", [6.0,2.0,1.0], rot=R)
study("elong-sc2", [6.0,2.0,1.0], sigma_cluster=2.0)
study("rod", [8.0,1.0,1.0])
