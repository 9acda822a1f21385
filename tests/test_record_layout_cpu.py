"""The per-subject records the CUDA step kernel reads, checked on the CPU: jmb_debug_pack_host runs the library's own
packer (jmb_create without a device) and a NumPy interpreter below reads the records with the index arithmetic of
jmb_step_kernel (jmbayes_b200/csrc/jmb_kernels.cuh: hazard-row blocks of RP lanes, band offsets, per-draw blocks,
longitudinal rows after the last block) and must reproduce the oracle's log_pyb / log_h / H / log_ptb.  No compute call
of the library is involved; this guards the layout for the model shapes the GPU tests cover, including several
transition rows per subject (multi-state fits).
"""
import ctypes as C

import numpy as np
import pytest

from jmbayes_b200 import _abi
from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort

TF = {"identity": lambda x: x, "expit": lambda x: np.exp(x) / (1 + np.exp(x)), "exp": np.exp}


def pack(cuda_lib, c, multi_state):
    L = cuda_lib.lib()
    vp = C.c_void_p
    L.jmb_debug_pack_host.argtypes = [C.POINTER(_abi.JmbData), C.POINTER(vp)]
    L.jmb_debug_layout.argtypes = [vp, _abi.c_int32_p, _abi.c_int64_p, _abi.c_int64_p]
    L.jmb_debug_records.argtypes = [vp, C.POINTER(_abi.JmbDraw), _abi.c_double_p, _abi.c_double_p]
    d, k1 = _abi.marshal_data(c["Data"], c["Covs"]["b"], c["interval_cens"], multi_state)
    w, k2 = _abi.marshal_draw(c["Data"], c["interval_cens"])
    h = vp()
    assert L.jmb_debug_pack_host(C.byref(d), C.byref(h)) == 0, L.jmb_last_error().decode()
    n = d.n
    ints = np.zeros(64, dtype=np.int32); doff = np.zeros(n + 1, dtype=np.int64); coff = np.zeros(n + 1, dtype=np.int64)
    assert L.jmb_debug_layout(h, ints.ctypes.data_as(_abi.c_int32_p), doff.ctypes.data_as(_abi.c_int64_p),
                              coff.ctypes.data_as(_abi.c_int64_p)) == 0
    drec = np.zeros(doff[n]); crec = np.zeros(coff[n])
    assert L.jmb_debug_records(h, C.byref(w), drec.ctypes.data_as(_abi.c_double_p), crec.ctypes.data_as(_abi.c_double_p)) == 0
    L.jmb_destroy(h)
    return ints, doff, coff, drec, crec


def interpret(c, ints, doff, coff, drec, crec, b, Bs, gam, alp):
    """What jmb_step_kernel (MODE_EVAL) computes, read from the records."""
    D = c["Data"]
    (RP, o_V, o_W2c, o_Pw, o_W1off, o_W1v, o_W2, o_long, c_long, HB, _ss, WB, _G, w2_const, MS, S) = [int(x) for x in ints[:16]]
    o_ZZ, o_U = ints[16:32], ints[32:48]
    T, O = len(D["XXbetas"]), len(D["y"])
    zz_ones0 = [int(ints[48 + t]) & 1 for t in range(T)]
    u_ones = [(int(ints[48 + t]) >> 1) & 1 for t in range(T)]
    z_ones0 = [(int(ints[56 + o]) >> 8) & 1 for o in range(O)]
    long_cols = [int(ints[56 + o]) >> 16 for o in range(O)]
    n, p = b.shape[0], len(gam)
    K = 15
    B = len(Bs)
    pyb = np.zeros(n); log_h, H, HL, ptb_rows, ptb_sub = [], [], [], [], np.zeros(n)
    for s in range(n):
        rec, cr = drec[doff[s]:doff[s + 1]], crec[coff[s]:coff[s + 1]]
        ntr = int(rec[0]) if MS else 1
        v = rec[o_V:o_V + (O + 1) // 2].view(np.int32)[:O]
        lrec, lcrec = o_long + (ntr - 1) * HB, c_long * ntr
        for k in range(O):
            vk = int(v[k]); cols = D["RE_inds"][k]
            y = rec[lrec:lrec + vk]
            eta = cr[lcrec:lcrec + vk].copy()
            if z_ones0[k]:
                eta += b[s, cols[0]]
            for j in range(z_ones0[k], len(cols)):
                eta += rec[lrec + (1 + j - z_ones0[k]) * vk: lrec + (2 + j - z_ones0[k]) * vk] * b[s, cols[j]]
            fam = D["fams"][k]
            if fam == "gaussian":
                ld = -0.5 * ((y - eta) / D["sigmas"][k]) ** 2
            elif fam == "binomial":
                pr = np.exp(eta) / (1 + np.exp(eta)); ld = y * np.log(pr) + (1 - y) * np.log(1 - pr)
            else:
                ld = y * eta - np.exp(eta)
            pyb[s] += ld.sum()
            lrec += long_cols[k] * vk; lcrec += vk
        for tb in range(ntr):
            hr, hc = rec[tb * HB:], cr[tb * c_long:]
            off = hr[o_W1off:o_W1off + RP // 2].view(np.int32)
            lin = np.zeros(RP)
            for r in range(RP):
                if r > K * S:
                    continue
                l1 = sum(hr[o_W1v + j * RP + r] * Bs[off[r] + j] for j in range(WB))
                l2 = sum((rec[o_W2c + j] if w2_const else hr[o_W2 + j * RP + r]) * gam[j] for j in range(p))
                l3 = 0.0
                for t in range(T):
                    cols = D["RE_inds2"][t]
                    z = b[s, cols[0]] if zz_ones0[t] else 0.0
                    for j in range(zz_ones0[t], len(cols)):
                        z += hr[o_ZZ[t] + (j - zz_ones0[t]) * RP + r] * b[s, cols[j]]
                    val = TF[D["trans_Funs"][t]](hc[t * RP + r] + z)
                    for u, col in enumerate(D["col_inds"][t]):
                        l3 += (1.0 if u_ones[t] else hr[o_U[t] + u * RP + r]) * val * alp[col]
                lin[r] = (l1 + l2) + l3
            Hs = float(np.sum(hr[o_Pw + 1:o_Pw + 1 + K] * np.exp(lin[1:1 + K])))
            HLs = float(np.sum(hr[o_Pw + 1 + K:o_Pw + 1 + 2 * K] * np.exp(lin[1 + K:1 + 2 * K]))) if S == 2 else Hs
            ev = hr[o_Pw] if MS else rec[0]
            if S == 2:
                t = (lin[0] if ev == 1 else 0.0) - (Hs if ev in (0.0, 1.0) else 0.0)
                if ev == 2: t += np.log(1 - np.exp(-Hs))
                if ev == 3: t += np.log(np.exp(-HLs) - np.exp(-Hs))
            else:
                t = ev * lin[0] - Hs
            log_h.append(lin[0]); H.append(Hs); HL.append(HLs); ptb_rows.append(t); ptb_sub[s] += t
    return dict(log_pyb=pyb, log_h=np.array(log_h), H=np.array(H), HL=np.array(HL), log_ptb=np.array(ptb_rows), ptb_sub=ptb_sub)


CASES = [("config2", False), ("config3", False), ("config4", False), ("config3", True), ("multistate", True)]


@pytest.mark.parametrize("cfg,ms", CASES)
def test_packed_records_reproduce_the_oracle_pieces(cuda_lib, oracle, cfg, ms):
    c = make_multistate_cohort(n=60) if cfg == "multistate" else make_cohort(cfg, n=60)
    D, ini = c["Data"], c["inits"]
    ints, doff, coff, drec, crec = pack(cuda_lib, c, ms)
    assert int(ints[14]) == int(ms)
    b = np.asarray(ini["b"]); Bs, gam, alp = (np.asarray(ini[k], dtype=float) for k in ("Bs_gammas", "gammas", "alphas"))
    got = interpret(c, ints, doff, coff, drec, crec, b, Bs, gam, alp)
    oracle.set_rowsum_mode(1)
    ref = oracle.eval_logpost_re(D, c["Covs"]["b"], c["interval_cens"], b, Bs, gam, alp, multiState=ms)
    for key in ("log_pyb", "log_h", "H", "log_ptb") + (("HL",) if c["interval_cens"] else ()):
        assert got[key].shape == ref[key].shape, key
        assert np.allclose(got[key], ref[key], rtol=1e-11, atol=1e-12), (key, float(np.max(np.abs(got[key] - ref[key]))))
    if ms:     # rowsum(log_ptb, idT_rsum): the b-step's per-subject survival term
        idT = np.asarray(D["idT2"][0])
        assert np.allclose(got["ptb_sub"], np.bincount(idT, weights=ref["log_ptb"]), rtol=1e-11, atol=1e-12)


def test_multistate_records_have_one_block_per_transition(cuda_lib):
    c = make_multistate_cohort(n=40)
    ints, doff, coff, drec, crec = pack(cuda_lib, c, True)
    idT = np.asarray(c["Data"]["idT2"][0]); ntr = np.bincount(idT)
    o_Pw, o_long, c_long, HB = int(ints[3]), int(ints[7]), int(ints[8]), int(ints[9])
    assert HB == o_long - o_Pw and int(ints[13]) == 0                     # w2_const is never set for these fits
    ev = np.asarray(c["Data"]["event"]); first = np.concatenate([[0], np.cumsum(ntr)])
    for s in range(40):
        rec = drec[doff[s]:doff[s + 1]]
        assert int(rec[0]) == ntr[s]
        for tb in range(ntr[s]):
            assert rec[tb * HB + o_Pw] == ev[first[s] + tb]               # the block's event indicator
        visits = sum(int(np.sum(np.asarray(i) == s)) for i in c["Data"]["idL"])
        assert coff[s + 1] - coff[s] >= c_long * ntr[s] + visits
