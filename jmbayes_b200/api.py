"""Host-side mirror of the reference's Rcpp interface for the mvJointModelBayes hot path.

The reference's callers are R functions; R is not available in this image, so this module plays
the role of the Rcpp shim for the tests and the benchmark: same entry-point names, same argument
lists (R lists become dicts with the reference's field names, see ``cohorts.py``), same returned
structure and the same error behaviour (a failing call raises, like ``Rcpp::stop``).  All compute
goes through the C ABI of ``libjmb200.so`` (``include/jmb200.h``); there is no CPU fallback - if
the CUDA library is missing or no GPU is usable, calls raise ``RuntimeError``.

reference: R/RcppExports.R:4-42, src/RcppExports.cpp:216-228, src/fast_LAPRWM.cpp:431-896.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from .build import LIB

_lib_cache = None


class JmbError(RuntimeError):
    pass


def lib():
    """Load ``libjmb200.so`` (fails loudly when it has not been built)."""
    global _lib_cache
    if _lib_cache is not None:
        return _lib_cache
    path = os.environ.get("JMB_LIB", LIB)      # A/B builds of the same library (experiments only)
    if not os.path.exists(path):
        raise JmbError(f"{path} is missing: build it with __graft_entry__.build() "
                       "(nvcc, sm_100a); there is no CPU fallback")
    _lib_cache = _bind(C.CDLL(path))
    return _lib_cache


def _bind(L):
    """Declare the argument types of the C ABI on a loaded library."""
    L.jmb_last_error.restype = C.c_char_p
    L.jmb_stream.restype = C.c_void_p
    L.jmb_kernel_name.restype = C.c_char_p
    L.jmb_kernel_name.argtypes = [C.c_void_p]
    L.jmb_stream.argtypes = [C.c_void_p]
    vp = C.c_void_p
    L.jmb_create.argtypes = [C.POINTER(_abi.JmbData), C.c_int, C.POINTER(vp)]
    L.jmb_destroy.argtypes = [vp]
    L.jmb_set_draw.argtypes = [vp, C.POINTER(_abi.JmbDraw)]
    L.jmb_set_xdesigns.argtypes = [vp, C.POINTER(_abi.JmbXDesigns)]
    L.jmb_set_draw_betas.argtypes = [vp, _abi.c_double_pp, _abi.c_double_p, _abi.c_double_p]
    L.jmb_lap_rwm.argtypes = [vp, C.POINTER(_abi.JmbPriors), C.POINTER(_abi.JmbControl),
                              C.POINTER(_abi.JmbChainIn), C.POINTER(_abi.JmbChainOut)]
    L.jmb_chain_begin.argtypes = [vp, C.POINTER(_abi.JmbPriors), C.POINTER(_abi.JmbControl),
                                  C.POINTER(_abi.JmbChainIn)]
    L.jmb_chain_iterate.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.jmb_chain_iterate_after.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.jmb_chain_end.argtypes = [vp, C.POINTER(_abi.JmbChainOut)]
    L.jmb_eval_logpost_re.argtypes = [vp, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p,
                                      _abi.c_double_p, C.POINTER(_abi.JmbEvalOut)]
    L.jmb_get_row_ptr.argtypes = [vp, C.c_int, _abi.c_int64_p]
    L.jmb_chain_slots.argtypes = [vp, C.c_int]
    L.jmb_select_chain.argtypes = [vp, C.c_int]
    L.jmb_chains_iterate.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.jmb_set_wlong.argtypes = [vp, _abi.c_double_p, _abi.c_double_p]
    post = [vp, C.c_double, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p, C.POINTER(_abi.JmbPriors),
            C.c_double, C.c_double, C.c_double, _abi.c_double_p]
    L.jmb_logPosterior.argtypes = post
    sv = [vp, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p]
    L.jmb_log_post_RE_svft.argtypes = sv
    L.jmb_survPred_svft_2.argtypes = sv
    L.jmb_lap_rwm_woRE.argtypes = [vp, C.POINTER(_abi.JmbPriors), C.POINTER(_abi.JmbControl), C.POINTER(_abi.JmbChainIn),
                                   C.POINTER(_abi.JmbChainOut), _abi.c_double_p]
    L.jmb_gradient_logPosterior.argtypes = post
    L.jmb_marglogLik2.argtypes = [vp, _abi.c_double_p, _abi.c_double_p, _abi.c_double_p, C.c_double, C.POINTER(_abi.JmbPriors),
                                  C.c_double, C.c_int, C.POINTER(_abi.JmbMllOut)]
    L.jmb_layout_info.argtypes = [vp, _abi.c_double_p, _abi.c_double_p, _abi.c_int32_p, _abi.c_int32_p]
    L.jmb_launch_count.argtypes = [vp, _abi.c_int64_p]
    L.jmb_launch_count_timed.argtypes = [vp, _abi.c_int64_p]
    L.jmb_comm_unique_id.argtypes = [C.c_char_p]
    L.jmb_comm_init.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.jmb_p2p_export.argtypes = [vp, C.c_int, C.c_char_p]
    L.jmb_p2p_attach.argtypes = [vp, C.c_int, C.c_int, C.c_char_p]
    L.jmb_device_count.argtypes = [C.POINTER(C.c_int)]
    return L


def _check(rc):
    if rc != 0:
        raise JmbError(lib().jmb_last_error().decode())


def device_count() -> int:
    c = C.c_int(0)
    _check(lib().jmb_device_count(C.byref(c)))
    return c.value


def _dptr(a):
    return a.ctypes.data_as(_abi.c_double_p)


class Fit:
    """The static designs of one ``mvJointModelBayes`` fit, resident on one GPU.

    Plays the role of the ``Data`` unpacking at src/fast_LAPRWM.cpp:434-505, done once per fit
    instead of once per chain.
    """

    def __init__(self, Data: dict, Covs_b, interval_cens: bool = False, multiState: bool = False,
                 device: int = 0, subject_offset: int = 0, basis_from_knots: bool = False):
        self._h = C.c_void_p()
        self.interval_cens = bool(interval_cens)
        d, keep = _abi.marshal_data(Data, Covs_b, interval_cens, multiState, subject_offset, basis_from_knots=basis_from_knots)
        self.n, self.q = d.n, d.q
        self.B, self.p, self.A = d.n_Bs, d.n_gammas, d.n_alphas
        self.O, self.T, self.K = d.n_outcomes, d.n_terms, d.K
        # multi-state fits: nT transition rows in the survival part, nRisks strata (src/fast_LAPRWM.cpp:519-529,607)
        self.multiState = bool(multiState)
        self.nT = d.nT if multiState else d.n
        self.nRisks = d.nRisks if multiState else 1
        _check(lib().jmb_create(C.byref(d), device, C.byref(self._h)))
        del keep

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if self._h:
            lib().jmb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- per-draw data (R/mvJointModelBayes.R:760-770) ------------------------------------------
    def set_draw(self, Data: dict):
        w, keep = _abi.marshal_draw(Data, self.interval_cens)
        _check(lib().jmb_set_draw(self._h, C.byref(w)))
        del keep

    def set_xdesigns(self, Data: dict):
        """Upload Data$X / XX / XXs[_int] once per fit so that Xbetas_calc (R/supportFuns_mvJointModelBayes.R:43-55)
        runs on the device for every draw (:meth:`set_draw_betas`)."""
        x, keep = _abi.marshal_xdesigns(Data, self.interval_cens)
        _check(lib().jmb_set_xdesigns(self._h, C.byref(x)))
        del keep

    def set_draw_betas(self, betas, sigmas, invD):
        """The per-draw step of runParallel (R/mvJointModelBayes.R:760-770) from the draw's ``betas`` (list per outcome),
        ``sigmas`` and ``invD``: a few hundred bytes host-to-device instead of the N-vectors of :meth:`set_draw`."""
        k = _abi.Keep()
        bl = k.f64_list([np.ascontiguousarray(b, dtype=np.float64) for b in betas])
        sig = [1.0 if (s is None or (isinstance(s, float) and np.isnan(s))) else float(s) for s in sigmas]
        _check(lib().jmb_set_draw_betas(self._h, bl, k.f64(sig), k.f64(invD)))
        del k

    # -- lap_rwm_C --------------------------------------------------------------------------------
    def _marshal_chain(self, initials, priors, scales, Covs, control, chain_id):
        pr, k1 = _abi.marshal_priors(priors)
        ct = _abi.marshal_control(control, chain_id)
        ci, k2 = _abi.marshal_chain_in(initials, scales, Covs)
        return pr, ct, ci, (k1, k2)

    def _result(self, bufs, n_out):
        mcmc = {k: bufs[k] for k in ("b", "Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas",
                                     "tau_Bs_gammas", "tau_gammas", "tau_alphas", "tau_td_alphas")}
        return dict(
            mcmc=mcmc, logWeights=bufs["logWeights"],
            scales=dict(sigma=dict(b=bufs["sigma_b"], Bs_gammas=float(bufs["sigma_surv"][0]),
                                   gammas=float(bufs["sigma_surv"][1]), alphas=float(bufs["sigma_surv"][2])),
                        Sigma=dict(Bs_gammas=bufs["Sigma_Bs_gammas"], gammas=bufs["Sigma_gammas"],
                                   alphas=bufs["Sigma_alphas"])),
            extras=dict(accept_b=bufs.get("accept_b"), accept_surv=bufs.get("accept_surv"),
                        trace_surv=bufs.get("trace_surv")))

    def lap_rwm_C(self, initials, priors, scales, Covs, control, chain_id: int = 0):
        """One chain for the current draw; returns the list ``lap_rwm_C`` returns
        (src/fast_LAPRWM.cpp:870-895) plus ``extras`` (acceptance rates, per-iteration sums)."""
        pr, ct, ci, keep = self._marshal_chain(initials, priors, scales, Covs, control, chain_id)
        total, n_out = _abi.chain_dims(control)
        out, bufs = _abi.alloc_chain_out(self.n, self.q, self.B, self.p, self.A, pr.n_td, n_out, total, n_risks=self.nRisks)
        _check(lib().jmb_lap_rwm(self._h, C.byref(pr), C.byref(ct), C.byref(ci), C.byref(out)))
        del keep
        return self._result(bufs, n_out)

    # -- stepwise driver (benchmark / tests) ----------------------------------------------------------
    def chain_begin(self, initials, priors, scales, Covs, control, chain_id: int = 0):
        pr, ct, ci, keep = self._marshal_chain(initials, priors, scales, Covs, control, chain_id)
        self._chain_meta = (pr.n_td,) + _abi.chain_dims(control)   # same control for every slot of a fit
        _check(lib().jmb_chain_begin(self._h, C.byref(pr), C.byref(ct), C.byref(ci)))
        del keep

    def chain_iterate(self, iters: int, time_kernels: bool = False, lead: int = 0):
        """Run ``iters`` iterations on the device; returns (loop ms, step-kernel ms or None).  ``lead`` untimed iterations are
        enqueued in the same call right before the timed ones (multi-rank runs: the clock starts after their exchange)."""
        ms, kms = C.c_float(0), C.c_float(0)
        if lead:
            _check(lib().jmb_chain_iterate_after(self._h, int(lead), int(iters), C.byref(ms)))
            return ms.value, None
        _check(lib().jmb_chain_iterate(self._h, int(iters), C.byref(ms),
                                       C.byref(kms) if time_kernels else None))
        return ms.value, (kms.value if time_kernels else None)

    # -- several chains (draws) of the fit in flight (R/mvJointModelBayes.R:860-905) ---------------------
    def chain_slots(self, count: int):
        _check(lib().jmb_chain_slots(self._h, int(count)))

    def select_chain(self, slot: int):
        """Slot that set_draw / chain_begin / chain_iterate / chain_end / lap_rwm_C act on."""
        _check(lib().jmb_select_chain(self._h, int(slot)))

    def chains_iterate(self, nslots: int, iters: int) -> float:
        """Advance slots 0..nslots-1 together by ``iters`` iterations; returns the device ms."""
        ms = C.c_float(0)
        _check(lib().jmb_chains_iterate(self._h, int(nslots), int(iters), C.byref(ms)))
        return ms.value

    def chain_end(self, fetch: bool = True):
        if not fetch:
            _check(lib().jmb_chain_end(self._h, None))
            return None
        n_td, total, n_out = self._chain_meta
        out, bufs = _abi.alloc_chain_out(self.n, self.q, self.B, self.p, self.A, n_td, n_out, total, n_risks=self.nRisks)
        _check(lib().jmb_chain_end(self._h, C.byref(out)))
        return self._result(bufs, n_out)

    # -- parity probes -------------------------------------------------------------------------------
    def eval_logpost_re(self, b, Bs_gammas, gammas, alphas):
        """Per-subject log p(y|b), log p(b), log h, H, HL, log p(T|b) at (b, params)
        (src/fast_LAPRWM.cpp:619-650); for a multi-state fit the last four have one value per transition row."""
        n = self.n
        b = np.asfortranarray(np.asarray(b, dtype=np.float64))
        res = {k: np.zeros(n if k in ("log_pyb", "log_pb") else self.nT) for k in ("log_pyb", "log_pb", "log_h", "H", "HL", "log_ptb")}
        out = _abi.JmbEvalOut()
        for k, a in res.items():
            setattr(out, k, _dptr(a))
        g = np.ascontiguousarray(np.atleast_1d(np.asarray(gammas, dtype=np.float64)))
        if g.size == 0:
            g = np.zeros(1)
        Bs = np.ascontiguousarray(np.asarray(Bs_gammas, dtype=np.float64))
        al = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
        _check(lib().jmb_eval_logpost_re(self._h, _dptr(b), _dptr(Bs), _dptr(g), _dptr(al), C.byref(out)))
        return res

    # -- survival-block Laplace step (marglogLik2's fn / gr, R/margLogLik2.R:42-65) --------------------
    def set_wlong(self, Wlong, Wlongs):
        """Data$Wlong (n x A) and Data$Wlongs (nK x A) for logPosterior / gradient_logPosterior."""
        a = np.asfortranarray(np.asarray(Wlong, dtype=np.float64))
        b = np.asfortranarray(np.asarray(Wlongs, dtype=np.float64))
        if a.shape != (self.nT, self.A) or b.shape != (self.nT * self.K, self.A):
            raise JmbError("set_wlong: Wlong must be n x n_alphas and Wlongs nK x n_alphas (nT rows for a multi-state fit)")
        _check(lib().jmb_set_wlong(self._h, _dptr(a), _dptr(b)))

    def _post_args(self, Bs_gammas, gammas, alphas, priors):
        g = np.ascontiguousarray(np.atleast_1d(np.asarray(gammas, dtype=np.float64)))
        if g.size == 0:
            g = np.zeros(1)
        Bs = np.ascontiguousarray(np.asarray(Bs_gammas, dtype=np.float64))
        al = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
        pr, keep = _abi.marshal_priors(priors)
        return Bs, g, al, pr, keep

    def logPosterior(self, temp, Bs_gammas, gammas, alphas, priors, tauBs, A_tauBs, B_tauBs) -> float:
        """logPosterior[_nogammas] (src/fast_logPost.cpp:8-46) on the designs of this fit."""
        Bs, g, al, pr, keep = self._post_args(Bs_gammas, gammas, alphas, priors)
        out = np.zeros(1)
        _check(lib().jmb_logPosterior(self._h, float(temp), _dptr(Bs), _dptr(g), _dptr(al), C.byref(pr),
                                      float(tauBs), float(A_tauBs), float(B_tauBs), _dptr(out)))
        return float(out[0])

    def gradient_logPosterior(self, temp, Bs_gammas, gammas, alphas, priors, tauBs, A_tauBs, B_tauBs):
        """gradient_logPosterior[_nogammas] (src/fast_score.cpp:8-84): score of (Bs, gammas, alphas, log tauBs)."""
        Bs, g, al, pr, keep = self._post_args(Bs_gammas, gammas, alphas, priors)
        out = np.zeros(self.B + self.p + self.A + 1)
        _check(lib().jmb_gradient_logPosterior(self._h, float(temp), _dptr(Bs), _dptr(g), _dptr(al), C.byref(pr),
                                               float(tauBs), float(A_tauBs), float(B_tauBs), _dptr(out)))
        return out

    def marglogLik2(self, thetas: dict, priors: dict, temp: float = 1.0, fixed_tau_Bs_gammas: bool = False):
        """marglogLik2(thetas, Data, priors, temp, fixed_tau_Bs_gammas) (R/margLogLik2.R) on the designs of this fit and
        the association designs of :meth:`set_wlong`: the BFGS driver runs natively in the library, every evaluation on
        the device.  Returns the value plus ``par``, ``hessian`` and - when tau is fixed - ``Covs`` / ``inits`` (the
        attributes of the reference's result).  ``thetas['tau_Bs_gammas']`` is on the log scale when it is not fixed."""
        Bs, g, al, pr, keep = self._post_args(thetas["Bs_gammas"], thetas.get("gammas", np.zeros(0)), thetas["alphas"], priors)
        P = self.B + self.p + self.A
        d = P if fixed_tau_Bs_gammas else P + 1
        tau = float(np.asarray(thetas["tau_Bs_gammas"]).ravel()[0])
        bufs = dict(par=np.zeros(d), hessian=np.zeros((d, d), order="F"), Cov_Bs_gammas=np.zeros((self.B, self.B), order="F"),
                    Cov_gammas=np.zeros((max(self.p, 1), max(self.p, 1)), order="F"), Cov_alphas=np.zeros((self.A, self.A), order="F"))
        out = _abi.JmbMllOut()
        for name, a in bufs.items():
            setattr(out, name, _dptr(a))
        _check(lib().jmb_marglogLik2(self._h, _dptr(Bs), _dptr(g), _dptr(al), tau, C.byref(pr), float(temp),
                                     int(bool(fixed_tau_Bs_gammas)), C.byref(out)))
        res = dict(value=out.value, opt_value=out.opt_value, convergence=out.convergence, counts=(out.counts[0], out.counts[1]),
                   par=bufs["par"], hessian=bufs["hessian"], restarts=int(out.restarts))
        if fixed_tau_Bs_gammas:
            res["Covs"] = dict(Bs_gammas=bufs["Cov_Bs_gammas"], gammas=bufs["Cov_gammas"][:self.p, :self.p], alphas=bufs["Cov_alphas"])
            res["inits"] = dict(Bs_gammas=bufs["par"][:self.B], gammas=bufs["par"][self.B:self.B + self.p],
                                alphas=bufs["par"][self.B + self.p:P], tau_Bs_gammas=tau)
        return res

    def lap_rwm_C_woRE(self, initials, priors, scales, Covs, control, chain_id: int = 0):
        """lap_rwm_C_woRE[_nogammas] (src/fast_LAPRWM.cpp:934-1297) on the designs of ``set_wlong``."""
        ini = dict(initials); ini.setdefault("b", np.zeros((1, 1)))
        sc = dict(scales); sc.setdefault("b", np.zeros(1))
        pr, k1 = _abi.marshal_priors(priors)
        ct = _abi.marshal_control(control, chain_id)
        ci, k2 = _abi.marshal_chain_in(ini, sc, Covs)
        total, n_out = _abi.chain_dims(control)
        out, bufs = _abi.alloc_chain_out(1, 1, self.B, self.p, self.A, 0, n_out, total)
        logPost = np.zeros(max(n_out, 1))
        _check(lib().jmb_lap_rwm_woRE(self._h, C.byref(pr), C.byref(ct), C.byref(ci), C.byref(out), _dptr(logPost)))
        mcmc = {k: bufs[k] for k in ("Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas", "tau_Bs_gammas",
                                     "tau_gammas", "tau_alphas")}
        return dict(mcmc=mcmc, logPost=logPost[:n_out],
                    scales=dict(sigma=dict(Bs_gammas=float(bufs["sigma_surv"][0]), gammas=float(bufs["sigma_surv"][1]),
                                           alphas=float(bufs["sigma_surv"][2])),
                                Sigma=dict(Bs_gammas=bufs["Sigma_Bs_gammas"], gammas=bufs["Sigma_gammas"],
                                           alphas=bufs["Sigma_alphas"])),
                    extras=dict(trace_surv=bufs["trace_surv"], accept_surv=bufs["accept_surv"]))

    def _svft(self, fn, b, Bs_gammas, gammas, alphas):
        b = np.asfortranarray(np.asarray(b, dtype=np.float64))
        g = np.ascontiguousarray(np.atleast_1d(np.asarray(gammas, dtype=np.float64)))
        if g.size == 0:
            g = np.zeros(1)
        Bs = np.ascontiguousarray(np.asarray(Bs_gammas, dtype=np.float64))
        al = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
        out = np.zeros(self.n)
        _check(fn(self._h, _dptr(b), _dptr(Bs), _dptr(g), _dptr(al), _dptr(out)))
        return out

    def log_post_RE_svft(self, b, Bs_gammas, gammas, alphas):
        """log_post_RE_svft (src/survfitJM_mvJMbayes.cpp:190-238) for every subject of the handle."""
        return self._svft(lib().jmb_log_post_RE_svft, b, Bs_gammas, gammas, alphas)

    def survPred_svft_2(self, b, Bs_gammas, gammas, alphas):
        """survPred_svft_2 (src/survfitJM_mvJMbayes.cpp:241-271) for every subject of the handle."""
        return self._svft(lib().jmb_survPred_svft_2, b, Bs_gammas, gammas, alphas)

    def row_ptr(self, outcome: int):
        rp = np.zeros(self.n + 1, dtype=np.int64)
        _check(lib().jmb_get_row_ptr(self._h, outcome, rp.ctypes.data_as(_abi.c_int64_p)))
        return rp

    def layout_info(self):
        sb, cb = C.c_double(0), C.c_double(0)
        wb, g = C.c_int32(0), C.c_int32(0)
        _check(lib().jmb_layout_info(self._h, C.byref(sb), C.byref(cb), C.byref(wb), C.byref(g)))
        return dict(shared_bytes_per_subject=sb.value, chain_bytes_per_subject=cb.value,
                    w1_band=wb.value, group_lanes=g.value,
                    kernel=lib().jmb_kernel_name(self._h).decode())

    def launch_count_timed(self) -> int:
        c = C.c_int64(0)
        _check(lib().jmb_launch_count_timed(self._h, C.byref(c)))
        return c.value

    def launch_count(self) -> int:
        c = C.c_int64(0)
        _check(lib().jmb_launch_count(self._h, C.byref(c)))
        return c.value

    # -- multi-GPU ---------------------------------------------------------------------------------
    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        _check(lib().jmb_comm_init(self._h, nranks, rank, unique_id))

    def p2p_export(self, nranks: int) -> bytes:
        buf = C.create_string_buffer(64)
        _check(lib().jmb_p2p_export(self._h, nranks, buf))
        return buf.raw

    def p2p_attach(self, nranks: int, rank: int, handles: bytes):
        _check(lib().jmb_p2p_attach(self._h, nranks, rank, handles))


def connect_ranks(fit: "Fit", world: int, rank: int, dist) -> None:
    """One process per GPU: NCCL communicator for the library plus the peer-memory mailboxes of the
    fused per-iteration all-reduce.  ``dist`` is an initialised ``torch.distributed`` module."""
    uid = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    fit.comm_init(world, rank, uid[0])
    handles = [None] * world
    dist.all_gather_object(handles, fit.p2p_export(world))
    fit.p2p_attach(world, rank, b"".join(handles))
    dist.barrier()


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().jmb_comm_unique_id(buf))
    return buf.raw


def lap_rwm_C(initials, Data, priors, scales, Covs, control, interval_cens=False, multiState=False,
              device: int = 0, chain_id: int = 0):
    """Same argument list as the reference's ``lap_rwm_C`` (R/RcppExports.R; src/fast_LAPRWM.cpp:431)."""
    with Fit(Data, Covs["b"], interval_cens, multiState, device) as fit:
        fit.set_draw(Data)
        return fit.lap_rwm_C(initials, priors, scales, Covs, control, chain_id)
