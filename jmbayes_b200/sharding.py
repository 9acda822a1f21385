"""Contiguous subject shards for one-process-per-GPU runs.

Each rank owns a contiguous block of subjects with all of their longitudinal rows, hazard rows,
random effects, proposal factors and scales; the survival parameters, priors and the random key
are replicated.  The only exchange per iteration is the sum over ranks of three scalars (sum of
log p(T|b) under the current and proposed survival parameters, and the log-weight sum), done inside the
finalize kernel over NVLink peer-memory mailboxes (ncclAllReduce only as the A/B switch
JMB_ALLREDUCE=nccl).  Draws are keyed by the global subject index
(``subject_offset + local index``), so a sharded run reproduces the single-GPU chain up to the
rounding of that three-scalar sum.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int, rows_per_subject=None):
    """Split ``n`` subjects into ``world`` contiguous blocks balanced by work (rows)."""
    if rows_per_subject is None:
        w = np.ones(n)
    else:
        w = np.asarray(rows_per_subject, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    targets = cum[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(cum, targets, side="left")
    cuts = np.clip(cuts, 1, n - 1) if n > 1 else cuts
    b = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    for r in range(1, world + 1):          # every rank gets at least one subject
        b[r] = max(b[r], b[r - 1] + 1) if r < world else n
    return b


def shard_cohort(c: dict, rank: int, world: int) -> dict:
    """Slice a cohort dict (``cohorts.make_cohort``) to the subjects of ``rank``."""
    D = c["Data"]
    ms = bool(c.get("multiState", False))
    # rows of the survival part: one per subject, or the transition rows of a multi-state fit (grouped by subject,
    # Data$idT2); a shard takes all the transition rows of its subjects
    idT = np.asarray(D["idT2"][0]) if ms else None
    n = int(np.asarray(c["inits"]["b"]).shape[0]) if ms else int(np.asarray(D["event"]).shape[0])
    nT = int(np.asarray(D["event"]).shape[0])
    K = int(np.asarray(D["Pw"]).shape[0]) // nT
    O = len(D["y"])
    counts = np.zeros(n)
    for k in range(O):
        counts += np.bincount(np.asarray(D["idL"][k]), minlength=n)
    ntr = np.bincount(idT, minlength=n) if ms else np.ones(n, dtype=np.int64)
    b = shard_bounds(n, world, counts + K * ntr)
    lo, hi = int(b[rank]), int(b[rank + 1])
    m = hi - lo
    t0, t1 = (int(np.searchsorted(idT, lo, "left")), int(np.searchsorted(idT, hi, "left"))) if ms else (lo, hi)
    qs = slice(t0 * K, t1 * K)

    def rows(a, sl):
        a = np.asarray(a)
        return np.asfortranarray(a[sl]) if a.ndim == 2 else a[sl].copy()

    out = dict(D)
    ys, Xb, Zs, idLs, idL2s = [], [], [], [], []
    for k in range(O):
        idL = np.asarray(D["idL"][k])
        r0, r1 = np.searchsorted(idL, lo, "left"), np.searchsorted(idL, hi, "left")
        ys.append(np.asarray(D["y"][k])[r0:r1].copy())
        Xb.append(np.asarray(D["Xbetas"][k])[r0:r1].copy())
        Zs.append(rows(D["Z"][k], slice(r0, r1)))
        idLs.append((idL[r0:r1] - lo).astype(np.int32))
        idL2s.append((np.asarray(D["idL2"][k])[lo:hi] - r0).astype(np.int32))
    out.update(y=ys, Xbetas=Xb, Z=Zs, idL=idLs, idL2=idL2s)
    sub = slice(t0, t1)
    for key in ("event", "W1", "W2", "Time", "strata", "TimeL"):
        if key in D:
            out[key] = rows(D[key], sub)
    for key in ("W1s", "W2s", "Pw"):
        out[key] = rows(D[key], qs)
    for key in ("XXbetas", "ZZ", "U"):
        out[key] = [rows(x, sub) for x in D[key]]
    for key in ("XXsbetas", "ZZs", "Us"):
        out[key] = [rows(x, qs) for x in D[key]]
    if c["interval_cens"]:
        for key in ("W1s_int", "W2s_int", "Pw_int"):
            out[key] = rows(D[key], qs)
        for key in ("XXsbetas_int", "ZZs_int", "Us_int"):
            out[key] = [rows(x, qs) for x in D[key]]
    out["idGK_fast"] = (np.arange(1, (t1 - t0) + 1) * K - 1).astype(np.int32)
    out["idT_rsum"] = np.arange(m, dtype=np.int32)
    if ms:
        loc = (idT[t0:t1] - lo).astype(np.int32)
        T = len(D["XXbetas"])
        out["idT"] = out["idT2"] = [loc for _ in range(T)]
        out["idTs"] = out["idT2s"] = [np.repeat(loc, K) for _ in range(T)]
        out["idT_rsum"] = np.flatnonzero(np.r_[loc[1:] != loc[:-1], True]).astype(np.int32)
    inits = dict(c["inits"]); inits["b"] = np.asfortranarray(np.asarray(c["inits"]["b"])[lo:hi])
    scales = dict(c["scales"]); scales["b"] = np.asarray(c["scales"]["b"])[lo:hi].copy()
    Covs = dict(c["Covs"]); Covs["b"] = np.asfortranarray(np.asarray(c["Covs"]["b"])[:, :, lo:hi])
    res = dict(c)
    res.update(Data=out, inits=inits, scales=scales, Covs=Covs, n=m, subject_offset=lo, bounds=b)
    return res
