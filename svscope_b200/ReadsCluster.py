"""Host-side mirror of src/ReadsCluster.py: same entry points, numeric work on the device.

``EMCluster`` keeps the reference's schedule and constants (K = 1..min(9, N-1), 20 EM steps,
clamp 1e-10, exp clip +-700, BIC with ZeroParamNum, the K=1 -> K=2 override) and its return
list ``[K, seqdatamx, Rclust, thetap, gamma, pie, BICList]`` (reference :221-277).  The Ward
tree and its cuts stay on host scipy exactly as in the reference (:243, :94)."""
from __future__ import annotations

import numpy as np

from . import batch as _batch
from ._lib import Context

np.random.seed(_batch.SEED)  # the reference seeds the global RNG at import (:42)


def CallDistance(read1, read2):
    """Fraction of equal positions (reference :44-50)."""
    total = len(read1) or 1
    return int((np.asarray(read1) == np.asarray(read2)).sum()) / total


def pariwiseDistance(seqdatamx):
    """N x N identity-fraction matrix with unit diagonal (reference :52-59), counted on the GPU."""
    X = np.asarray(seqdatamx)
    N, nf = X.shape
    if N == 0:
        return np.eye(0)
    if nf == 0:
        return np.eye(N)
    ctx = Context.default()
    _, _, _, ident = _batch.msa_features(ctx, [X], [np.zeros(nf, np.uint8)], [-1.0])[0]
    sim = ident.astype(np.float64) / nf
    np.fill_diagonal(sim, 1.0)
    return sim


def CheckParam(AimParam, episilon=1e-10):
    return np.clip(AimParam, episilon, 1 - episilon)


def BIC(ParamDict, ZeroParamNum=0):
    """2*loglik - n_theta*log(N) from the last iteration (reference :211-219)."""
    theta_f, pi_f, lik = ParamDict["theta"][-1], ParamDict["pi"][-1], ParamDict["likelihood"][-1]
    n_theta = len(pi_f) - 1 + theta_f.shape[0] * theta_f.shape[1] * (theta_f.shape[2] - 1) - ZeroParamNum
    return 2 * lik.sum() - n_theta * np.log(len(lik))


def EM(K, seqdatamx, initselection=1, Nstep=20, Z=None):
    """One mixture fit (reference :190-209).  Returns the last state in the reference's
    ParamDict layout (lists with the final element only: intermediate states are not kept)."""
    from scipy.cluster.hierarchy import fcluster
    if initselection != 1 or Z is None:
        raise NotImplementedError("only the hierarchical initialisation (initselection=1) is on the path")
    ctx = Context.default()
    X = np.asarray(seqdatamx)
    labels = (fcluster(Z, K, criterion="maxclust") - 1).astype(np.int32)
    res = _batch.em_batch(ctx, [X], [_batch.EmTaskSpec(0, K, labels, n_steps=Nstep)], want_theta=True)[0]
    if res["status"] >= 0:
        res, _ = _batch._resume_fit(ctx, X, K, labels, res, True)
    return {"pi": [res["pi"]], "theta": [res["theta"]], "gamma": [res["gamma"]], "likelihood": [res["lik"]]}


def EMCluster(seqdatamx, initselection=1, max_C=9, ShowPlot=False):
    """Cluster reads with the categorical mixture model; reference :221-277."""
    if initselection != 1:
        raise NotImplementedError("only initselection=1 (hierarchical initialisation) is on the path")
    if not 2 <= int(max_C) <= _batch.MAX_C:
        raise NotImplementedError("max_C from 2 to 9 (the mixture kernel is built for K <= 9; the reference passes 9)")
    ctx = Context.default()
    X = np.asarray(seqdatamx)
    nf = X.shape[1]
    _, _, zp, ident = _batch.msa_features(ctx, [X], [np.zeros(nf, np.uint8)], [-1.0])[0]
    sim = ident.astype(np.float64) / (nf if nf else 1)
    np.fill_diagonal(sim, 1.0)
    # no per-call reseed here: like the reference, EMCluster draws from the process-wide RNG
    fit = _batch.em_cluster_many(ctx, [X], [sim], [zp], want_theta=True, reseed=False, max_C=int(max_C))[0]
    return [fit["K"], seqdatamx, fit["labels"], fit["theta"], fit["gamma"], fit["pi"], fit["bics"]]
