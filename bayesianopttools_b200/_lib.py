"""ctypes binding of libkrigb200.so (include/krigb200.h).

There is NO CPU fallback: if the shared library is missing, or no sm_100
device is visible when a compute entry point is called, this raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# KRIGB200_LIB: development aid, an A/B build of the same library (bayesianopttools_b200/build.py --variant=)
LIB_PATH = os.environ.get("KRIGB200_LIB") or os.path.join(_HERE, "csrc", "libkrigb200.so")

KERNEL_IDS = {"gaussian": 0, "iso_gaussian": 0, "exponential": 1, "matern32": 2, "matern52": 3}
OUT_IDS = {"pred": 0, "ssqr": 1, "s": 2, "ei": 3, "poi": 4, "pof": 5, "lcb": 6, "ebe": 7, "u": 8,
           "fpc": 9}
OUT_COUNT = 10
NORM_DEFAULT, NORM_STD, NORM_NONE = 0, 1, 2
REFINE_AUTO, REFINE_OFF, REFINE_ON = -1, 0, 1
YMAP_NONE, YMAP_DEFAULT, YMAP_STD = 0, 1, 2
TYPE_KRIGING, TYPE_KPLS = 0, 1
ERR_NOT_PD = 4

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class AcqParams(C.Structure):
    _fields_ = [("ybest", C.c_double), ("limit", C.c_double), ("limit_sign", C.c_int),
                ("sigmalcb", C.c_double)]


class CmaesOpts(C.Structure):
    _fields_ = [("popsize", C.c_int), ("maxiter", C.c_int), ("check_every", C.c_int), ("sigma0", C.c_double),
                ("tolfun", C.c_double), ("tolx", C.c_double), ("seed", C.c_ulonglong)]


class SweepResult(C.Structure):
    _fields_ = [("best_val", C.c_double), ("best_idx", C.c_longlong), ("min_u", C.c_double),
                ("min_u_idx", C.c_longlong), ("n_fail", C.c_longlong), ("n_fail_minus", C.c_longlong),
                ("n_fail_plus", C.c_longlong), ("n_points", C.c_longlong)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class KrigB200Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("krigb200 status %d: %s" % (status, msg))
        self.status = status


_lib = None

# every symbol include/krigb200.h declares: (name, restype, argtypes)
_SIGNATURES = [
    ("kb_version", C.c_int, []),
    ("kb_last_error", C.c_char_p, []),
    ("kb_device_count", C.c_int, []),
    ("kb_model_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, c_double_p,
                                  c_double_p, c_double_p, C.c_int, c_int_p, C.c_int, C.c_int, c_double_p]),
    ("kb_model_destroy", None, [C.c_void_p]),
    ("kb_model_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("kb_model_use_own_stream", C.c_int, [C.c_void_p]),
    ("kb_model_sync", C.c_int, [C.c_void_p]),
    ("kb_nll_batch", C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_double, c_double_p,
                               c_int_p]),
    ("kb_nll_batch_dev", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                   C.c_void_p, C.c_void_p]),
    ("kb_nll_batch_extras", C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    ("kb_reserve_population", C.c_int, [C.c_void_p, C.c_int]),
    ("kb_fit_state", C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_double, c_double_p,
                               c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    ("kb_model_set_trend", C.c_int, [C.c_void_p, c_int_p, C.c_int, C.c_int]),
    ("kb_model_set_norm", C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p]),
    ("kb_model_set_predict_refine", C.c_int, [C.c_void_p, C.c_int]),
    ("kb_model_predict_refined", C.c_int, [C.c_void_p]),
    ("kb_model_set_ymap", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    ("kb_predict", C.c_int, [C.c_void_p, C.c_longlong, c_double_p, C.POINTER(AcqParams),
                             C.POINTER(c_double_p), C.c_int, C.POINTER(SweepResult)]),
    ("kb_predict_dev", C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                 C.POINTER(AcqParams), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    ("kb_points_create", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_int, c_double_p]),
    ("kb_points_destroy", None, [C.c_void_p]),
    ("kb_points_generate", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_double,
                                     C.c_double, c_double_p, c_double_p, C.c_ulonglong, C.c_longlong]),
    ("kb_points_generate_lowdisc", C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_longlong, C.c_int, C.c_int,
                                             C.POINTER(C.c_uint), c_double_p, c_double_p, C.c_longlong]),
    ("kb_points_get_rows", C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, c_double_p]),
    ("kb_predict_points", C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(AcqParams), C.POINTER(c_double_p), C.c_int,
                                    C.POINTER(SweepResult)]),
    ("kb_exi2d", C.c_int, [C.c_int, C.c_longlong, c_double_p, c_double_p, C.c_int, c_double_p, c_double_p,
                           c_double_p]),
    ("kb_ehvi2d", C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, c_double_p, C.c_int, c_double_p, c_double_p,
                            C.c_int, c_double_p, C.POINTER(C.c_longlong), c_double_p]),
    ("kb_ehvi2d_points", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_int,
                                   c_double_p, C.POINTER(C.c_longlong), c_double_p]),
    ("kb_train_cmaes", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, c_double_p, c_double_p, c_double_p,
                                 c_double_p, C.c_void_p, c_double_p, c_double_p, c_int_p, c_int_p, c_double_p]),
    ("kb_loocv", C.c_int, [C.c_void_p, c_double_p]),
    ("kb_corr_matrix", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p,
                                 C.c_int, C.c_int, c_double_p, c_double_p]),
    ("kb_measure_peak", C.c_int, [C.c_int, C.c_int, c_double_p]),
    ("kb_set_profiling", C.c_int, [C.c_void_p, C.c_int]),
    ("kb_get_stage_ms", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    ("kb_launch_count", C.c_longlong, []),
]
EXPORTED_SYMBOLS = [s[0] for s in _SIGNATURES]


def load():
    """Load the shared library (build it first with bayesianopttools_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libkrigb200.so is not built (%s). Run `python -m bayesianopttools_b200.build`; "
            "the krig-b200 hot path has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, res, args in _SIGNATURES:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise KrigB200Error(status, load().kb_last_error().decode("utf-8", "replace"))


def require_gpu():
    if load().kb_device_count() < 1:
        raise RuntimeError("krig-b200: no CUDA device visible; this path runs on sm_100a only "
                           "(there is no CPU fallback)")


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def default_device():
    return int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("KRIGB200_USE_LOCAL_RANK", "1") == "1" \
        else 0


class DeviceModel:
    """Owner of one ``kb_model`` handle (device copies of X_norm, y, F + workspaces)."""

    def __init__(self, X_norm, y, F, kernels, model_type="kriging", plscoeff=None, device=None):
        require_gpu()
        lib = load()
        self.X_norm = _f64(X_norm)
        self.y = _f64(np.asarray(y).reshape(-1))
        self.n, self.d = self.X_norm.shape
        self.F = _f64(np.asarray(F).reshape(self.n, -1))
        self.nind = self.F.shape[1]
        self.kernels = [k.lower() for k in kernels]
        for k in self.kernels:
            if k not in KERNEL_IDS:
                raise ValueError("Kernel option not available!")
        ids = np.array([KERNEL_IDS[k] for k in self.kernels], dtype=np.int32)
        self.model_type = TYPE_KPLS if model_type.lower() == "kpls" else TYPE_KRIGING
        h, pls = 0, None
        if self.model_type == TYPE_KPLS:
            pls = _f64(plscoeff)
            h = pls.shape[1]
        self.h = h
        self.device = default_device() if device is None else int(device)
        hdl = C.c_void_p()
        check(lib.kb_model_create(C.byref(hdl), self.device, self.n, self.d, self.nind, _dp(self.X_norm),
                                  _dp(self.y), _dp(self.F), len(ids), ids.ctypes.data_as(c_int_p),
                                  self.model_type, h, _dp(pls) if pls is not None else None))
        self._h = hdl
        self._lib = lib
        self.fitted = False

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- population likelihood ---------------------------------------------
    def nll_batch(self, xpop, trainvar, fixed_nugget, return_info=False):
        xpop = _f64(np.atleast_2d(xpop))
        B, nbhyp = xpop.shape
        nll = np.empty(B)
        info = np.zeros(B, dtype=np.int32)
        check(self._lib.kb_nll_batch(self._h, B, nbhyp, _dp(xpop), int(bool(trainvar)),
                                     float(fixed_nugget), _dp(nll), info.ctypes.data_as(c_int_p)))
        return (nll, info) if return_info else nll

    def nll_batch_dev(self, xpop_ptr, B, nbhyp, trainvar, fixed_nugget, nll_ptr, info_ptr=None):
        check(self._lib.kb_nll_batch_dev(self._h, B, nbhyp, C.c_void_p(xpop_ptr), int(bool(trainvar)),
                                         float(fixed_nugget), C.c_void_p(nll_ptr),
                                         C.c_void_p(info_ptr) if info_ptr else None))

    def nll_extras(self, B):
        beta = np.empty((B, self.nind))
        s2 = np.empty(B)
        check(self._lib.kb_nll_batch_extras(self._h, B, _dp(beta), _dp(s2)))
        return beta, s2

    def train_cmaes(self, x0, sigma0, trainvar, fixed_nugget, lb=None, ub=None, scaling=None, popsize=0, maxiter=0,
                    seed=0, tolfun=0.0, tolx=0.0, check_every=0):
        """CMA-ES over the hyper-parameters with the generation loop on the device (kb_train_cmaes).
        Returns (xbest, fbest, generations, stop_reason, history of the best objective per generation)."""
        x0 = _f64(np.asarray(x0).reshape(-1))
        N = x0.size
        lam = int(popsize) if popsize else 4 + int(3 * np.log(N))
        mi = int(maxiter) if maxiter else 100 + 150 * (N + 3) ** 2 // int(np.sqrt(lam))
        arr = lambda v: None if v is None else _f64(np.broadcast_to(np.asarray(v, dtype=float), (N,)))
        lb_, ub_, sc_ = arr(lb), arr(ub), arr(scaling)
        opts = CmaesOpts(int(popsize), int(maxiter), int(check_every), float(sigma0), float(tolfun), float(tolx), int(seed))
        xbest, hist = np.empty(N), np.zeros(mi)
        fbest, gens, why = C.c_double(), C.c_int(), C.c_int()
        check(self._lib.kb_train_cmaes(self._h, N, int(bool(trainvar)), float(fixed_nugget), _dp(x0),
                                       _dp(lb_) if lb_ is not None else None, _dp(ub_) if ub_ is not None else None,
                                       _dp(sc_) if sc_ is not None else None, C.byref(opts), _dp(xbest),
                                       C.cast(C.byref(fbest), c_double_p), C.cast(C.byref(gens), c_int_p),
                                       C.cast(C.byref(why), c_int_p), _dp(hist)))
        return xbest, float(fbest.value), int(gens.value), int(why.value), hist[:gens.value].copy()

    def reserve_population(self, B):
        check(self._lib.kb_reserve_population(self._h, int(B)))

    def set_profiling(self, on=True):
        check(self._lib.kb_set_profiling(self._h, int(bool(on))))

    def stage_ms(self, back=0):
        """(decode, build, chol, gls, predict) device milliseconds, `back` calls ago (-1: none)."""
        buf = (C.c_float * 5)()
        check(self._lib.kb_get_stage_ms(self._h, int(back), buf))
        return tuple(float(v) for v in buf)

    def set_stream(self, cuda_stream):
        """Run on a caller-owned CUDA stream (an integer handle); None returns to the handle's own stream."""
        if cuda_stream is None:
            check(self._lib.kb_model_use_own_stream(self._h))
        else:
            check(self._lib.kb_model_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def sync(self):
        check(self._lib.kb_model_sync(self._h))

    # -- fit ------------------------------------------------------------------
    def fit_state(self, x, trainvar, fixed_nugget, want_U=True, want_Psi=True):
        x = _f64(np.atleast_1d(x))
        U = np.empty((self.n, self.n)) if want_U else None
        Psi = np.empty((self.n, self.n)) if want_Psi else None
        BE = np.empty(self.nind)
        s2, nll, nug = C.c_double(), C.c_double(), C.c_double()
        wg = np.empty(len(self.kernels))
        check(self._lib.kb_fit_state(self._h, _dp(x), len(x), int(bool(trainvar)), float(fixed_nugget),
                                     _dp(U) if want_U else None, _dp(Psi) if want_Psi else None, _dp(BE),
                                     C.byref(s2), C.byref(nll), C.byref(nug), _dp(wg)))
        self.fitted = True
        return dict(U=U, Psi=Psi, BE=BE.reshape(-1, 1), SigmaSqr=s2.value, NegLnLike=nll.value,
                    nugget=nug.value, wgkf=wg)

    def set_predict_refine(self, mode):
        """Corrected sweeps on ill-conditioned fits: 'auto' (default: fits whose pivots reach down to the nugget),
        False / 0 (always the plain explicit-inverse product) or True / 1.  Takes effect at the next fit_state."""
        code = {"auto": REFINE_AUTO, None: REFINE_AUTO, False: REFINE_OFF, True: REFINE_ON, 0: REFINE_OFF, 1: REFINE_ON,
                -1: REFINE_AUTO}[mode]
        check(self._lib.kb_model_set_predict_refine(self._h, code))

    @property
    def predict_refined(self):
        """Whether the last fit_state chose the corrected sweeps."""
        return bool(self._lib.kb_model_predict_refined(self._h))

    def set_trend(self, idx):
        """Trend multi-index (nind x d).  The reference-facing layer calls this before every prediction;
        an unchanged table is not uploaded again."""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        key = (idx.shape, idx.tobytes())
        if getattr(self, "_trend_key", None) == key:
            return
        check(self._lib.kb_model_set_trend(self._h, idx.ctypes.data_as(c_int_p), idx.shape[0], idx.shape[1]))
        self._trend_key = key

    def set_norm(self, norm_type, p0=None, p1=None):
        """Input normalisation of the predictor (unchanged parameters are not uploaded again)."""
        if norm_type == NORM_NONE:
            key = (NORM_NONE,)
            if getattr(self, "_norm_key", None) == key:
                return
            check(self._lib.kb_model_set_norm(self._h, NORM_NONE, None, None))
        else:
            p0, p1 = _f64(np.asarray(p0).reshape(-1)), _f64(np.asarray(p1).reshape(-1))
            key = (norm_type, p0.tobytes(), p1.tobytes())
            if getattr(self, "_norm_key", None) == key:
                return
            check(self._lib.kb_model_set_norm(self._h, norm_type, _dp(p0), _dp(p1)))
        self._norm_key = key

    def set_ymap(self, mode, a=1.0, b=0.0):
        """Map of the predicted mean to real units for models trained on normalised responses (prediction.py:213-217)."""
        key = (int(mode), float(a), float(b))
        if getattr(self, "_ymap_key", (YMAP_NONE, 1.0, 0.0)) == key:
            return
        check(self._lib.kb_model_set_ymap(self._h, *key))
        self._ymap_key = key

    def loocv(self):
        pred = np.empty(self.n)
        check(self._lib.kb_loocv(self._h, _dp(pred)))
        return pred

    # -- predict --------------------------------------------------------------
    def predict(self, x, outputs=("pred",), acq="pred", ybest=0.0, limit=0.0, limit_sign=1, sigmalcb=0.0,
                want_result=True):
        x = _f64(np.atleast_2d(x))
        m = x.shape[0]
        if x.shape[1] != self.d:
            raise ValueError("prediction sites have %d columns, model has %d" % (x.shape[1], self.d))
        ptrs = (c_double_p * OUT_COUNT)()
        outs = {}
        for name in outputs:
            buf = np.empty(m)
            outs[name] = buf
            ptrs[OUT_IDS[name]] = _dp(buf)
        ap = AcqParams(float(ybest), float(limit), int(limit_sign), float(sigmalcb))
        res = SweepResult()
        check(self._lib.kb_predict(self._h, m, _dp(x), C.byref(ap), ptrs, OUT_IDS[acq],
                                   C.byref(res) if want_result else None))
        return outs, (res.as_dict() if want_result else None)

    def predict_points(self, pts, outputs=(), acq="pred", ybest=0.0, limit=0.0, limit_sign=1, sigmalcb=0.0):
        """Sweep a DevicePoints set: same outputs / reductions as predict(), no host->device copy."""
        if pts.d != self.d:
            raise ValueError("point set has %d columns, model has %d" % (pts.d, self.d))
        ptrs = (c_double_p * OUT_COUNT)()
        outs = {}
        for name in outputs:
            buf = np.empty(pts.n)
            outs[name] = buf
            ptrs[OUT_IDS[name]] = _dp(buf)
        ap = AcqParams(float(ybest), float(limit), int(limit_sign), float(sigmalcb))
        res = SweepResult()
        check(self._lib.kb_predict_points(self._h, pts._h, C.byref(ap), ptrs, OUT_IDS[acq], C.byref(res)))
        return outs, res.as_dict()

    def predict_dev(self, x_ptr, m, idx_offset=0, out_ptrs=None, acq="pred", ybest=0.0, limit=0.0,
                    limit_sign=1, sigmalcb=0.0, result_ptr=None):
        ptrs = (C.c_void_p * OUT_COUNT)()
        for name, p in (out_ptrs or {}).items():
            ptrs[OUT_IDS[name]] = p
        ap = AcqParams(float(ybest), float(limit), int(limit_sign), float(sigmalcb))
        check(self._lib.kb_predict_dev(self._h, int(m), C.c_void_p(x_ptr), int(idx_offset), C.byref(ap),
                                       ptrs, OUT_IDS[acq], C.c_void_p(result_ptr) if result_ptr else None))


class DevicePoints:
    """A point set resident on the device (kb_points): uploaded once, swept many times -- the fixed
    Monte-Carlo population of AK-MCS (akmcs.py:66-206)."""

    DISTS = {"normal": 0, "gaussian": 0, "lognormal": 1, "gumbel": 2, "random": 3, "uniform": 3, "halton": 4,
             "sobol": 5, "sobolnew": 5}

    def __init__(self, x, device=None):
        require_gpu()
        self._lib = load()
        x = _f64(np.atleast_2d(x))
        self.n, self.d = x.shape
        self.device = default_device() if device is None else int(device)
        h = C.c_void_p()
        check(self._lib.kb_points_create(C.byref(h), self.device, self.n, self.d, _dp(x)))
        self._h = h

    @classmethod
    def generate(cls, n, d, dist="gaussian", mean=0.0, stddev=1.0, lb=None, ub=None, seed=0, first_index=0,
                 device=None):
        """Point set generated on the device: random variates from the counter-based Philox4x32-10
        generator (reproducible per site index), or the 'halton' / 'sobol' designs of misc/sampling
        (bit-identical to the host samplers; site i is the sequence's point first_index + i), mapped to
        [lb, ub] when bounds are given."""
        require_gpu()
        self = cls.__new__(cls)
        self._lib = load()
        self.n, self.d = int(n), int(d)
        self.device = default_device() if device is None else int(device)
        kind = cls.DISTS[dist.lower()]
        lbp = ubp = None
        if kind >= 4:
            if (lb is None) != (ub is None):
                raise ValueError("give both bounds or neither")
            if lb is not None:
                lb_, ub_ = _f64(np.asarray(lb).reshape(-1)), _f64(np.asarray(ub).reshape(-1))
                lbp, ubp = _dp(lb_), _dp(ub_)
            dirs = None
            if kind == 5:
                from .misc.sampling.samplingplan import sobol_direction_numbers
                dirs_ = np.ascontiguousarray(sobol_direction_numbers(self.d, 32), dtype=np.uint32)
                dirs = dirs_.ctypes.data_as(C.POINTER(C.c_uint))
            h = C.c_void_p()
            check(self._lib.kb_points_generate_lowdisc(C.byref(h), self.device, self.n, self.d, kind, dirs, lbp, ubp,
                                                       int(first_index)))
            self._h = h
            return self
        if kind == 3:
            if lb is None or ub is None:
                raise ValueError("type 'random' is selected, please input lower bound and upper bound value")
            lb_, ub_ = _f64(np.asarray(lb).reshape(-1)), _f64(np.asarray(ub).reshape(-1))
            lbp, ubp = _dp(lb_), _dp(ub_)
        h = C.c_void_p()
        check(self._lib.kb_points_generate(C.byref(h), self.device, self.n, self.d, kind, float(mean), float(stddev),
                                           lbp, ubp, int(seed), int(first_index)))
        self._h = h
        return self

    def rows(self, row0, nrows=1):
        out = np.empty((int(nrows), self.d))
        check(self._lib.kb_points_get_rows(self._h, int(row0), int(nrows), _dp(out)))
        return out

    def __getitem__(self, key):
        """pts[i, :] -> one site as a (d,) array (what AKMCS needs for the next infill point)."""
        i = key[0] if isinstance(key, tuple) else key
        return self.rows(int(i), 1)[0]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.kb_points_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def corr_matrix(XN, XM, theta, kernel="gaussian", plscoeff=None, device=None):
    require_gpu()
    lib = load()
    XN, XM, theta = _f64(np.atleast_2d(XN)), _f64(np.atleast_2d(XM)), _f64(np.asarray(theta).reshape(-1))
    if kernel.lower() not in KERNEL_IDS:
        raise ValueError("Kernel option not available!")
    out = np.empty((XN.shape[0], XM.shape[0]))
    pls, h = None, 0
    if plscoeff is not None:
        pls = _f64(plscoeff)
        h = pls.shape[1]
    check(lib.kb_corr_matrix(default_device() if device is None else device, XN.shape[0], XM.shape[0],
                             XN.shape[1], _dp(XN), _dp(XM), _dp(theta), KERNEL_IDS[kernel.lower()], h,
                             _dp(pls) if pls is not None else None, _dp(out)))
    return out


def measure_peak(which, device=None):
    require_gpu()
    v = C.c_double()
    check(load().kb_measure_peak(default_device() if device is None else device, int(which), C.byref(v)))
    return v.value


def launch_count():
    return int(load().kb_launch_count())


def exi2d_batch(front, ref, mu, s, device=None):
    """`exi2d(P, r, mu, s)` (optim_tools/ehvi/exi2d.py:6-83) for an (m, 2) batch of candidates."""
    require_gpu()
    lib = load()
    front = _f64(np.atleast_2d(front))
    ref = _f64(np.asarray(ref).reshape(-1))
    mu = _f64(np.atleast_2d(mu))
    s = _f64(np.atleast_2d(s))
    if front.shape[1] != 2 or ref.size != 2 or mu.shape[1] != 2 or s.shape != mu.shape:
        raise ValueError("exi2d handles two objectives: P (k, 2), r (2,), mu and s (m, 2)")
    out = np.empty(mu.shape[0])
    check(lib.kb_exi2d(default_device() if device is None else int(device), mu.shape[0], _dp(mu), _dp(s),
                       front.shape[0], _dp(front), _dp(ref), _dp(out)))
    return out


def ehvi2d_models(model1, model2, x, front, ref, spread="ssqr", want_values=True):
    """Batched `ehvicalc` (optim_tools/ehvi/EHVIcomputation.py:10-42): both fitted DeviceModels predict
    the raw candidates x ((m, d) host array or a DevicePoints set); returns (improvement per candidate or None, argmax index, max value)."""
    front = _f64(np.atleast_2d(front))
    ref = _f64(np.asarray(ref).reshape(-1))
    if front.shape[1] != 2 or ref.size != 2:
        raise ValueError("ehvi2d handles two objectives: P (k, 2), r (2,)")
    bi, bv = C.c_longlong(-1), C.c_double(0.0)
    sp = 0 if spread.lower() == "ssqr" else 1
    if isinstance(x, DevicePoints):
        out = np.empty(x.n) if want_values else None
        check(model1._lib.kb_ehvi2d_points(model1._h, model2._h, x._h, front.shape[0], _dp(front), _dp(ref), sp,
                                           _dp(out) if out is not None else None, C.byref(bi),
                                           C.cast(C.byref(bv), c_double_p)))
        return out, int(bi.value), float(bv.value)
    x = _f64(np.atleast_2d(x))
    out = np.empty(x.shape[0]) if want_values else None
    check(model1._lib.kb_ehvi2d(model1._h, model2._h, x.shape[0], _dp(x), front.shape[0], _dp(front), _dp(ref),
                                sp, _dp(out) if out is not None else None,
                                C.byref(bi), C.cast(C.byref(bv), c_double_p)))
    return out, int(bi.value), float(bv.value)
