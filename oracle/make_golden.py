"""Generate tests/golden/*.npz by running the UNMODIFIED reference in-process.

TEST INFRASTRUCTURE ONLY.  Run in the build container (``/root/reference``
mounted):  ``python oracle/make_golden.py``.  The fixtures are committed; the
GPU box never runs this script.

Each fixture stores the INPUTS (so tests need neither the reference nor its
samplers) and the reference's OUTPUTS for the hot path:
``likelihood`` (likelihood_func.py:6), ``calckernel`` (kernelfunc.py:4),
``prediction`` (prediction.py:67), ``Kriging.train`` (kriging_model.py:169),
``loocv2`` (krigloocv2.py:7) and the AK-MCS reductions (akmcs.py:208-238).
The two notebook cases also carry the values STORED in the notebooks
(Kriging_Demo.ipynb:341-345, Single_Objective_Bayesian_Optimization.ipynb:307-310)
so the reference-as-run-here is itself pinned to its published outputs.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ref = load_reference()


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


def styb(X):
    return 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1, keepdims=True)


def make_info(X, y, lb, ub, kernel=("gaussian",), nugget=-6, order=0, nrestart=5,
              optimizer="lbfgsb", **extra):
    info = ref.initkriginfo("single")
    info.update(X=X.copy(), y=y.copy(), lb=np.asarray(lb, float), ub=np.asarray(ub, float),
                nrestart=nrestart, kernel=list(kernel), nugget=nugget, TrendOrder=order,
                optimizer=optimizer)
    info.update(extra)
    return info


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote", name, {k: np.shape(v) for k, v in arrs.items()})


def pred_all(obj, xq, types=("pred", "SSqr", "s", "EI", "poi")):
    with quiet():
        outs = obj.predict(xq, list(types))
    return {("p_" + t.lower()): np.asarray(o, float).ravel() for t, o in zip(types, outs)}


# ---------------------------------------------------------------------------
# 1/2. notebook cases: trained by the reference's own L-BFGS-B path
# ---------------------------------------------------------------------------
def notebook_case(name, nsample, stored):
    lb, ub = np.array([-5.0, -5.0]), np.array([5.0, 5.0])
    with quiet():
        _, X = ref.sampling("halton", 2, nsample, result="real", upbound=ub, lobound=lb)
        y = ref.evaluate_analytic(X, "styblinski")
        info = make_info(X, y, lb, ub)
        obj = ref.Kriging(info, standardization=True, standtype="default", normy=False,
                          trainvar=False)
        obj.train(disp=False)
        K = obj.KrigInfo
        loocv_err, loocv_pred = obj.loocvcalc(metrictype="rmse")
    xx = np.linspace(-5, 5, 100)
    gx, gy = np.meshgrid(xx, xx)
    grid = np.column_stack([gx.reshape(-1), gy.reshape(-1)])
    with quiet():
        ygrid = np.asarray(obj.predict(grid, ["pred"]), float).ravel()
    yact = styb(grid).ravel()
    rmse_grid = float(np.sqrt(np.mean((yact - ygrid) ** 2)))
    rng = np.random.default_rng(7)
    xq = rng.uniform(-5, 5, size=(257, 2))
    # start points of the reference's train() (Sobol in the hyper-parameter box)
    with quiet():
        _, xhyp = ref.sampling("sobol", 2, 5, result="real", upbound=K["ubhyp"], lobound=K["lbhyp"])
    save(name, X=X, y=y, lb=lb, ub=ub, X_norm=K["X_norm"], F=K["F"],
         theta_log10=K["Theta"], nll=K["NegLnLike"], BE=K["BE"], sigma2=np.asarray(K["SigmaSqr"]),
         U=K["U"], Psi=K["Psi"], loocv_rmse=loocv_err, loocv_pred=loocv_pred,
         grid_pred_sub=ygrid[::37], grid_idx=np.arange(10000)[::37], rmse_grid=rmse_grid,
         xq=xq, xhyp=xhyp, stored=np.array(stored, float), **pred_all(obj, xq))


# ---------------------------------------------------------------------------
# 3. randomised differential cases at the likelihood / prediction level
# ---------------------------------------------------------------------------
def random_case(name, n, d, kernel, trainvar=False, nugget=-6, order=0, seed=0, npop=6,
                kpls_h=None, predict=True, theta_lo=-0.5, theta_hi=0.8):
    rng = np.random.default_rng(seed)
    lb, ub = -5.0 * np.ones(d), 5.0 * np.ones(d)
    X = rng.uniform(-5, 5, size=(n, d))
    y = styb(X[:, :min(d, 10)])
    if d > 10:
        y = y + 0.1 * np.sum(X[:, 10:] ** 2, axis=1, keepdims=True)
    extra = {}
    with quiet():
        info = make_info(X, y, lb, ub, kernel=kernel, nugget=nugget, order=order, **extra)
        if kpls_h:
            info["n_princomp"] = kpls_h
            obj = ref.KPLS(info, standardization=True, standtype="default", normy=False,
                           trainvar=trainvar)
        else:
            obj = ref.Kriging(info, standardization=True, standtype="default", normy=False,
                              trainvar=trainvar)
    K = obj.KrigInfo
    nb = len(K["ubhyp"])
    nth = kpls_h if kpls_h else d
    xpop = np.empty((npop, nb))
    xpop[:, :nth] = rng.uniform(theta_lo, theta_hi, size=(npop, nth))
    pos = nth
    if trainvar:
        xpop[:, pos] = rng.uniform(3.0, 4.5, size=npop)      # log10 sigma^2 near var(y)
        pos += 1
    if isinstance(nugget, list):
        xpop[:, pos] = rng.uniform(nugget[0], nugget[1], size=npop)
        pos += 1
    if len(kernel) > 1:
        xpop[:, pos:pos + len(kernel)] = rng.uniform(0.05, 1.0, size=(npop, len(kernel)))
    nlls, BEs, s2s = [], [], []
    with quiet():
        for x in xpop:
            Kb = ref.likelihood(x.copy(), K, mode="all", trainvar=trainvar)
            nlls.append(float(Kb["NegLnLike"]))
            BEs.append(np.asarray(Kb["BE"], float).ravel())
            s2s.append(float(np.asarray(Kb["SigmaSqr"]).ravel()[0]))
    arrs = dict(X=X, y=y, lb=lb, ub=ub, X_norm=K["X_norm"], F=K["F"], idx=np.asarray(K["idx"]),
                xpop=xpop, nll=np.array(nlls), BE=np.array(BEs), sigma2=np.array(s2s),
                kernel=np.array(list(kernel)), trainvar=np.array(trainvar),
                nugget=np.array(nugget, float), order=np.array(order),
                U_last=Kb["U"], Psi_last=Kb["Psi"])
    if kpls_h:
        arrs["plscoeff"] = np.asarray(K["plscoeff"], float)
        arrs["n_princomp"] = np.array(kpls_h)
    if predict and order == 0:
        xq = rng.uniform(-5, 5, size=(133, d))
        xq[0] = X[3]                                      # an exact training site (s -> ~0)
        K["limit"], K["limittype"] = float(np.median(y)), ">="
        obj.KrigInfo = K
        arrs["xq"] = xq
        arrs["limit"] = np.array(K["limit"])
        arrs.update(pred_all(obj, xq, ("pred", "SSqr", "s", "EI", "poi", "pof")))
    save(name, **arrs)


def kernel_case():
    rng = np.random.default_rng(11)
    XN, XM = rng.uniform(-1, 1, (37, 4)), rng.uniform(-1, 1, (53, 4))
    theta = 10.0 ** rng.uniform(-0.5, 0.5, 4)
    pls = rng.normal(size=(4, 2))
    theta_h = 10.0 ** rng.uniform(-0.5, 0.5, 2)
    arrs = dict(XN=XN, XM=XM, theta=theta, pls=pls, theta_h=theta_h)
    for k in ("gaussian", "exponential", "matern32", "matern52"):
        arrs["K_" + k] = ref.calckernel(XN, XM, theta, 4, type=k)
    arrs["K_kpls"] = ref.calckernel(XN, XM, theta_h, 4, type="gaussian", plscoeff=pls)
    save("calckernel", **arrs)


def trend_case():
    arrs = {}
    rng = np.random.default_rng(5)
    for order, nvar in ((0, 3), (1, 4), (2, 3), (2, 5), (3, 2)):
        with quiet():
            idx = ref.polytruncation(order, nvar, 1)
            Xs = rng.uniform(-1, 1, (9, nvar))
            bound = np.vstack((-np.ones((1, nvar)), np.ones((1, nvar))))
            Fm = ref.compute_regression_mat(idx, Xs, bound, np.ones(nvar))
        arrs["idx_%d_%d" % (order, nvar)] = np.asarray(idx)
        arrs["Xs_%d_%d" % (order, nvar)] = Xs
        arrs["F_%d_%d" % (order, nvar)] = Fm
    save("trend", **arrs)


def sampling_case():
    arrs = {}
    with quiet():
        for nv, ns in ((2, 5), (10, 7), (3, 16)):
            arrs["sobol_%d_%d" % (nv, ns)] = ref.sampling("sobol", nv, ns)[0]
            arrs["halton_%d_%d" % (nv, ns)] = ref.sampling("halton", nv, ns)[0]
    save("sampling", **arrs)


def akmcs_case():
    """Four-branch system (testcase/RA/testcase.py:48-55): population stats of iteration 0."""
    rng = np.random.default_rng(3)
    lb, ub = np.array([-5.0, -5.0]), np.array([5.0, 5.0])
    with quiet():
        _, X = ref.sampling("halton", 2, 12, result="real", upbound=ub, lobound=lb)
        y = ref.evaluate_ra(X, "fourbranches")
        info = make_info(X, y, lb, ub, nrestart=5)
        obj = ref.Kriging(info, standardization=True, standtype="default", normy=False,
                          trainvar=False)
        obj.train(disp=False)
    pop = rng.standard_normal((4000, 2))
    with quiet():
        ak = ref.AKMCS(obj, dict(init_samp=pop, maxupdate=1, problem="fourbranches"))
        ak.run(autoupdate=False, disp=False, logging=False)
    K = obj.KrigInfo
    temp1 = ak.Gx - 1.96 * ak.sigmaG.reshape(-1, 1)
    temp2 = ak.Gx + 1.96 * ak.sigmaG.reshape(-1, 1)
    save("akmcs_fourbranch", X=X, y=y, lb=lb, ub=ub, X_norm=K["X_norm"], F=K["F"],
         theta_log10=K["Theta"], nll=K["NegLnLike"], pop=pop, Gx=ak.Gx.ravel(),
         sigmaG=ak.sigmaG.ravel(), Pf=ak.Pf, minU=ak.minU, argminU=int(np.argmin(ak.U)),
         xnew=ak.xnew, stop=ak.stop_criteria, n_fail=int(np.sum(ak.Gx <= 0)),
         n_fail_minus=int(np.sum(temp1 <= 0)), n_fail_plus=int(np.sum(temp2 <= 0)))

def sobol_indices(ref_root="/root/reference"):
    """tests/golden/sobol_indices.npz: the reference's SobolIndices on the analytical Styblinski-Tang
    function (no surrogate), nvar = 3, bounds [-5, 5] repeated for the 2 nvar sampling dimensions as
    the reference requires (sobol_ind.py:24-27 samples nvar*2 columns with the caller's bounds)."""
    import os as _os
    from oracle.ref_loader import load_reference
    load_reference()
    cwd = _os.getcwd()
    _os.chdir(_os.path.join(ref_root, "demo"))
    try:
        from sensitivity_analysis.sobol_ind import SobolIndices as RefSI
        lb, ub = -5 * np.ones(6), 5 * np.ones(6)
        si = RefSI(3, krigobj=None, problem="styblinski", ub=ub, lb=lb)
        res = si.analyze(True, True, True)
    finally:
        _os.chdir(cwd)
    np.savez(_os.path.join(OUT, "sobol_indices.npz"), lb=lb, ub=ub, A_head=si.A[:64], B_head=si.B[:64],
             A_tail=si.A[-64:], B_tail=si.B[-64:], first=res["first"], total=res["total"],
             second_keys=np.array(list(res["second"].keys())), second_vals=np.array(list(res["second"].values())))


def ehvi2d(ref_root="/root/reference"):
    """tests/golden/ehvi2d.npz: the reference's exi2d (optim_tools/ehvi/exi2d.py) on random 2-objective
    fronts of 1, 2, 5 and 9 points, 12 candidates each."""
    import os as _os
    from oracle.ref_loader import load_reference
    load_reference()
    cwd = _os.getcwd()
    _os.chdir(_os.path.join(ref_root, "demo"))
    try:
        from optim_tools.ehvi.exi2d import exi2d as ref_exi2d
        rng = np.random.default_rng(0)
        out = {}
        for c, k in enumerate((1, 2, 5, 9)):
            f1 = np.sort(rng.uniform(0, 1, k))
            f2 = np.sort(rng.uniform(0, 1, k))[::-1]
            P = np.column_stack([f1, f2])[rng.permutation(k)]
            r = np.array([1.3, 1.2])
            mu = rng.uniform(-0.2, 1.4, (12, 2))
            s = rng.uniform(0.01, 0.5, (12, 2))
            ref = np.array([ref_exi2d(P, r, mu[i], s[i]) for i in range(12)])
            for nm, v in zip(("P", "r", "mu", "s", "ref"), (P, r, mu, s, ref)):
                out["%s%d" % (nm, c)] = v
    finally:
        _os.chdir(cwd)
    np.savez(_os.path.join(OUT, "ehvi2d.npz"), **out)


def mobo_host(ref_root="/root/reference"):
    """tests/golden/mobo_host.npz: host-side pieces of the reference's multi-objective loop --
    `paretopoint` (optim_tools/searchpareto.py) on random and duplicated rows, and the two-objective test
    problems (testcase/analyticalfcn/cases.py).  (`paregopre` cannot be recorded: the reference's `scale`
    raises under numpy 2.x, samplingplan.py:141.)"""
    import os as _os
    from oracle.ref_loader import load_reference
    load_reference()
    cwd = _os.getcwd()
    _os.chdir(_os.path.join(ref_root, "demo"))
    try:
        from optim_tools.searchpareto import paretopoint as ref_pp
        from testcase.analyticalfcn.cases import evaluate as ref_eval
        rng = np.random.default_rng(21)
        out = {}
        for c, m in enumerate((1, 2, 9, 40)):
            B = rng.uniform(0, 1, (m, 2))
            if m == 9:
                B[5] = B[2]                      # duplicated row: the reference drops both copies
                B[7, 0] = B[1, 0]                # tie in one objective
            A, idx = ref_pp(B)
            out["B%d" % c], out["A%d" % c], out["I%d" % c] = B, np.asarray(A), np.asarray(idx)
        B3 = rng.uniform(0, 1, (25, 3))
        A, idx = ref_pp(B3)
        out["B4"], out["A4"], out["I4"] = B3, np.asarray(A), np.asarray(idx)
        X = rng.uniform(-1, 1, (7, 2))
        out["X"] = X
        for t in ("schaffer", "fonseca", "schaffer1"):
            out["y_" + t] = np.asarray(ref_eval(X, t))
            out["y1_" + t] = np.asarray(ref_eval(X[0], t))
    finally:
        _os.chdir(cwd)
    np.savez(_os.path.join(OUT, "mobo_host.npz"), **out)


def branch_cases():
    """tests/golden/branches.npz: branches of the prediction / likelihood boundary the other fixtures do not
    reach -- standtype 'std' (prediction.py:160), norm_y = True (prediction.py:213-217; Kriging.standardize itself
    crashes with normy=True in the reference -- kwarg mismatch at kriging_model.py:121-129 -- so the dict is
    completed by hand with the reference's own `standardize(..., norm_y=True)` and then goes through the
    reference's likelihood and prediction), limittype '<' (prediction.py:279-282), 'lcb' / 'ebe' outputs
    (:244-247) and the isotropic Gaussian kernel (likelihood_func.py:60-61)."""
    rng = np.random.default_rng(31)
    n, d = 45, 3
    lb, ub = -5.0 * np.ones(d), 5.0 * np.ones(d)
    X = rng.uniform(-5, 5, (n, d))
    y = styb(X)
    xq = rng.uniform(-5, 5, (61, d))
    x = np.array([0.1, 0.25, 0.3])
    types = ("pred", "SSqr", "s", "EI", "poi", "pof", "ebe", "fpc")
    out = dict(X=X, y=y, lb=lb, ub=ub, xq=xq, theta_log10=x)

    def run(tag, K, obj, xx=x):
        with quiet():
            K = ref.likelihood(xx.copy(), K, mode="all", trainvar=False)
        K["limit"], K["limittype"], K["sigmalcb"] = float(np.median(y)), "<", 2.0
        obj.KrigInfo = K
        out[tag + "_nll"] = np.array(K["NegLnLike"])
        out[tag + "_limit"] = np.array(K["limit"])
        for t, o in pred_all(obj, xq, types).items():
            out[tag + "_" + t] = o
        with quiet():
            lcb = np.asarray(obj.predict(xq, ["lcb"]), float)      # (m, m): f_i - sigmalcb SSqr_j
        out[tag + "_p_lcb_diag"] = np.diag(lcb).copy()
        return K

    with quiet():
        obj = ref.Kriging(make_info(X, y, lb, ub), standardization=True, standtype="std", normy=False, trainvar=False)
    K = run("std", obj.KrigInfo, obj)
    for k in ("X_norm", "X_mean", "X_std", "F"):
        out["std_" + k] = np.asarray(K[k], float)
    # norm_y = True, 'std': the reference's standardize returns the six arrays its Kriging.standardize fails to unpack
    with quiet():
        obj = ref.Kriging(make_info(X, y, lb, ub), standardization=True, standtype="std", normy=False, trainvar=False)
        Xn, yn, Xm, ym, Xs, ys = ref.standardize(X, y, type="std", norm_y=True)
    K = obj.KrigInfo
    K.update(X_norm=Xn, y_norm=yn, X_mean=Xm, y_mean=ym, X_std=Xs, y_std=ys, norm_y=True)
    K = run("stdny", K, obj)
    for k in ("X_norm", "y_norm", "X_mean", "X_std", "y_mean", "y_std", "F"):
        out["stdny_" + k] = np.asarray(K[k], float)
    # norm_y = True, 'default'
    with quiet():
        obj = ref.Kriging(make_info(X, y, lb, ub), standardization=True, standtype="default", normy=False, trainvar=False)
        rngs = np.vstack((np.hstack((lb, np.min(y))), np.hstack((ub, np.max(y)))))
        Xn, yn = ref.standardize(X, y, type="default", norm_y=True, range=rngs)
    K = obj.KrigInfo
    K.update(X_norm=Xn, y_norm=yn, norm_y=True)
    K = run("defny", K, obj)
    for k in ("X_norm", "y_norm", "F"):
        out["defny_" + k] = np.asarray(K[k], float)
    # isotropic Gaussian kernel: one length-scale for every dimension
    with quiet():
        obj = ref.Kriging(make_info(X, y, lb, ub, kernel=["iso_gaussian"]), standardization=True, standtype="default",
                          normy=False, trainvar=False)
        xi = np.array([0.2])
        out["iso_nll_default"] = np.array(ref.likelihood(xi.copy(), obj.KrigInfo, mode="default", trainvar=False))
    K = run("iso", obj.KrigInfo, obj, xx=xi)
    out["iso_X_norm"], out["iso_F"] = np.asarray(K["X_norm"], float), np.asarray(K["F"], float)
    save("branches", **out)


if __name__ == "__main__":
    notebook_case("krigdemo_nb", 50, [-0.06768278, -0.06558988, 121.47648552358743,
                                      4.379695408530209, 4.298317119160842])
    notebook_case("sobo_nb", 10, [-1.25, -1.25, 42.76734625432907, 80.00548066291594, np.nan])
    kernel_case()
    trend_case()
    sampling_case()
    random_case("ok_gaussian", 60, 3, ["gaussian"], seed=1)
    random_case("ok_exponential", 50, 3, ["exponential"], seed=2)
    random_case("ok_matern32", 50, 4, ["matern32"], seed=3)
    random_case("ok_matern52", 50, 4, ["matern52"], seed=4)
    random_case("ok_gaussian_trainvar", 40, 3, ["gaussian"], trainvar=True, seed=5)
    random_case("ok_gaussian_tunednugget", 40, 3, ["gaussian"], nugget=[-8, -3], seed=6)
    random_case("ok_composite", 45, 3, ["gaussian", "matern32"], seed=7)
    random_case("ok_composite_tn_tv", 45, 3, ["gaussian", "matern52"], trainvar=True,
                nugget=[-8, -3], seed=8)
    random_case("trend1", 70, 4, ["gaussian"], order=1, seed=9, predict=False)
    random_case("trend2", 90, 4, ["gaussian"], order=2, seed=10, predict=False)
    random_case("kpls", 80, 12, ["gaussian"], kpls_h=3, seed=12)
    random_case("ok_gaussian_n200_d10", 200, 10, ["gaussian"], seed=13, npop=4, theta_lo=0.0,
                theta_hi=0.6)
    akmcs_case()
    sobol_indices()
    ehvi2d()
    mobo_host()
    branch_cases()
