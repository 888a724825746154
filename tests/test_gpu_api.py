"""GPU: the reference-facing Python surface (Kriging / KPLS / likelihood / prediction / SOBO /
AKMCS) reproduces the reference's stored outputs end to end."""
import os

import numpy as np
import pytest

from conftest import assert_rel, load_golden, relerr

pytestmark = pytest.mark.gpu


def _info(g, **extra):
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    K = initkriginfo("single")
    K.update(X=g["X"].copy(), y=g["y"].copy(), lb=g["lb"].copy(), ub=g["ub"].copy(), nrestart=5,
             kernel=["gaussian"], nugget=-6, TrendOrder=0, optimizer="lbfgsb", problem="styblinski")
    K.update(extra)
    return K


@pytest.mark.parametrize("name", ["krigdemo_nb", "sobo_nb"])
def test_train_reproduces_notebook(name, capsys):
    """Kriging_Demo.ipynb:341-345 / Single_Objective_...ipynb:307-310: same Sobol starts, same
    L-BFGS-B (eps=1e-3 forward differences, evaluated as one batched stencil) -> same optimum."""
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    g = load_golden(name)
    obj = Kriging(_info(g), standardization=True, standtype="default", normy=False, trainvar=False)
    assert np.array_equal(obj.KrigInfo["X_norm"], g["X_norm"])
    obj.train(disp=False)
    K = obj.KrigInfo
    assert abs(K["NegLnLike"] - float(g["nll"])) / abs(float(g["nll"])) < 1e-9
    assert np.allclose(K["Theta"], g["theta_log10"], atol=2e-5)
    assert K["U"].shape == g["U"].shape and K["BE"].shape == (1, 1) and K["SigmaSqr"].shape == (1, 1)
    err, pred = obj.loocvcalc(metrictype="rmse")
    assert pred.shape == (len(g["y"]), 1)
    assert abs(err - float(g["loocv_rmse"])) / float(g["loocv_rmse"]) < 1e-5     # (theta* of the two L-BFGS-B runs differ by 1e-6)
    xx = np.linspace(-5, 5, 100)
    gx, gy = np.meshgrid(xx, xx)
    grid = np.column_stack([gx.reshape(-1), gy.reshape(-1)])
    ygrid = obj.predict(grid, ["pred"])
    assert ygrid.shape == (10000, 1)
    assert_rel("test_gpu_api.py:43 ygrid.ravel()[g['grid_idx']]", ygrid.ravel()[g["grid_idx"]], g["grid_pred_sub"], 1e-5)
    yact = 0.5 * np.sum(grid ** 4 - 16 * grid ** 2 + 5 * grid, axis=1)
    rmse = np.sqrt(np.mean((yact - ygrid.ravel()) ** 2))
    assert abs(rmse - float(g["rmse_grid"])) / float(g["rmse_grid"]) < 1e-5


def test_likelihood_and_prediction_shims_follow_reference_conventions():
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood, likelihood_batch
    from bayesianopttools_b200.surrogate_models.supports.kernelfunc import calckernel
    g = load_golden("ok_composite_tn_tv")
    K = _info(g, kernel=["gaussian", "matern52"], nugget=[-8, -3])
    obj = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=True)
    assert len(obj.KrigInfo["ubhyp"]) == g["xpop"].shape[1]
    for b in (0, 3):
        v = likelihood(g["xpop"][b], obj.KrigInfo, mode="default", trainvar=True)
        assert isinstance(v, float) and abs(v - g["nll"][b]) / abs(g["nll"][b]) < 1e-9
    vb = likelihood_batch(g["xpop"], obj.KrigInfo, trainvar=True)
    assert np.max(np.abs(vb - g["nll"]) / np.abs(g["nll"])) < 1e-9
    Kout = likelihood(g["xpop"][-1], obj.KrigInfo, mode="all", trainvar=True)
    assert Kout is obj.KrigInfo
    assert relerr(Kout["U"], g["U_last"]) < 1e-9 and relerr(Kout["Psi"], g["Psi_last"]) < 1e-13
    assert np.isscalar(Kout["SigmaSqr"]) or np.ndim(Kout["SigmaSqr"]) == 0
    with pytest.raises(TypeError):
        likelihood(g["xpop"][0], obj.KrigInfo, mode="bogus")
    # predict conventions (prediction.py:292-298)
    obj.KrigInfo["limit"], obj.KrigInfo["limittype"] = float(g["limit"]), ">="
    outs = obj.predict(g["xq"], ["pred", "SSqr", "s", "EI", "poi", "pof"])
    assert isinstance(outs, list) and outs[0].shape == (133, 1) and outs[1].shape == (1, 133)
    for o, k in zip(outs, ("pred", "ssqr", "s", "ei", "poi", "pof")):
        assert_rel("shim predict " + k, o, g["p_" + k], 1e-9)
    one = obj.predict(g["xq"][5], "EI")
    assert isinstance(one, float) and abs(one - g["p_ei"][5]) <= 1e-9 * np.max(np.abs(g["p_ei"]))
    with pytest.raises(ValueError):
        obj.predict(g["xq"][:2], ["nope"])
    with pytest.raises(ValueError):
        calckernel(g["X_norm"], g["X_norm"], np.ones(3), 3, type="cubic")
    k = calckernel(g["X_norm"][:5], g["X_norm"][:7], np.ones(3), 3, type="matern32")
    assert k.shape == (5, 7)


def test_likelihood_raises_linalgerror_when_not_pd():
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.likelihood_func import likelihood
    g = load_golden("ok_gaussian")
    X = g["X"].copy()
    X[1] = X[0]
    K = _info(g, X=X, nugget=-30)
    obj = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    with pytest.raises(np.linalg.LinAlgError):
        likelihood(np.zeros(3), obj.KrigInfo, mode="default", trainvar=False)
    with pytest.raises(np.linalg.LinAlgError):
        likelihood(np.zeros(3), obj.KrigInfo, mode="all", trainvar=False)


def test_kpls_train_cmaes_and_predict():
    from bayesianopttools_b200.surrogate_models.kpls_model import KPLS
    from oracle import kriging_oracle as ko
    g = load_golden("kpls")
    K = _info(g, n_princomp=3, optimizer="cmaes", nrestart=1, cmaes_popsize=16, cmaes_maxiter=40, seed=3)
    obj = KPLS(K, standardization=True, standtype="default", normy=False, trainvar=False)
    assert_rel("test_gpu_api.py:104 obj.KrigInfo['plscoeff'] ** 2", obj.KrigInfo["plscoeff"] ** 2, g["plscoeff"] ** 2, 1e-9)
    obj.train(disp=False)
    Kt = obj.KrigInfo
    ref = ko.nll(Kt["Theta"], g["X_norm"], g["y"], g["F"], plscoeff=Kt["plscoeff"], n_princomp=3, full=True)
    assert abs(Kt["NegLnLike"] - ref["NegLnLike"]) / abs(ref["NegLnLike"]) < 1e-9
    assert Kt["NegLnLike"] <= np.min(g["nll"]) + 1e-6          # at least as good as the random population
    pr = obj.predict(g["xq"], ["pred", "s"])
    out = ko.predict(ko.standardize_x(g["xq"], g["lb"], g["ub"]), g["X_norm"], g["y"], g["F"], g["idx"], ref,
                     plscoeff=Kt["plscoeff"])
    assert_rel("kpls pred", pr[0], out["pred"], 1e-9)
    assert_rel("kpls s", pr[1], out["s"], 1e-9)


def test_akmcs_iteration0_matches_reference():
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.reliability_analysis.akmcs import AKMCS
    g = load_golden("akmcs_fourbranch")
    obj = Kriging(_info(g, problem="fourbranches"), standardization=True, standtype="default", normy=False,
                  trainvar=False)
    obj.train(disp=False)
    assert abs(obj.KrigInfo["NegLnLike"] - float(g["nll"])) / abs(float(g["nll"])) < 1e-7
    ak = AKMCS(obj, dict(init_samp=g["pop"], maxupdate=2, problem="fourbranches"))
    ak.run(autoupdate=False, disp=False)
    assert ak.Gx.shape == (4000, 1) and ak.sigmaG.shape == (1, 4000)
    assert_rel("test_gpu_api.py:127 ak.Gx", ak.Gx, g["Gx"], 1e-6)
    assert ak.Pf == float(g["Pf"])
    assert np.array_equal(ak.xnew, g["xnew"])
    assert abs(ak.stop_criteria - float(g["stop"])) < 1e-12
    ak.run(autoupdate=True, disp=False)                          # two enrichment steps run end to end
    assert obj.KrigInfo["X"].shape[0] == len(g["X"]) + 2 and ak.updateX.shape == (3, 2)


def test_sobo_sweep_improves_best():
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.optim_tools.SOBO import SOBO
    g = load_golden("sobo_nb")
    obj = Kriging(_info(g, nrestart=3), standardization=True, standtype="default", normy=False, trainvar=False)
    obj.train(disp=False)
    y0 = float(np.min(obj.KrigInfo["y"]))
    sobo = SOBO(dict(nup=3, nrestart=2, acquifunc="EI", acquifuncopt="sweep", ncandidates=200000, nsamp=10,
                     rng=np.random.default_rng(0)), obj)
    sobo.run(disp=False)
    assert obj.KrigInfo["X"].shape[0] == 13
    assert float(np.min(obj.KrigInfo["y"])) <= y0


def test_sobol_indices_on_surrogate_match_the_oracle_and_the_function():
    """Sobol indices from the batched predictor.  The estimators evaluated
    on the device predictor equal the same estimators evaluated on the oracle's predictions (same
    Monte-Carlo matrices), and agree with the indices of the analytical function within the
    surrogate/Monte-Carlo error."""
    from oracle import kriging_oracle as ko
    from bayesianopttools_b200.sensitivity_analysis.sobol_ind import SobolIndices
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    rng = np.random.default_rng(5)
    nvar, n = 3, 150
    lb, ub = -5.0 * np.ones(nvar), 5.0 * np.ones(nvar)
    X = rng.uniform(-5, 5, (n, nvar))
    y = (50 - 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1)).reshape(-1, 1)
    K = initkriginfo("single")
    K.update(X=X, y=y, lb=lb, ub=ub, nrestart=3, kernel=["gaussian"], nugget=-6, TrendOrder=0,
             optimizer="lbfgsb", problem="styblinski")
    obj = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
    obj.train(disp=False)
    si = SobolIndices(nvar, krigobj=obj, ub=np.hstack([ub, ub]), lb=np.hstack([lb, lb]))
    res = si.analyze(True, True, True)
    # same estimators on the oracle's predictions of the same matrices
    KI = obj.KrigInfo
    fit = ko.nll(np.asarray(KI["Theta"]).ravel(), KI["X_norm"], KI["y"], KI["F"], full=True)
    pred = lambda M: ko.predict(ko.standardize_x(M, lb, ub), KI["X_norm"], KI["y"], KI["F"],
                                np.zeros((1, nvar), int), fit)["pred"].reshape(-1, 1)
    ya, yb = pred(si.A), pred(si.B)
    assert relerr(si.ya, ya) < 1e-8 and relerr(si.yb, yb) < 1e-8
    fo2 = (np.sum(ya) / si.n) ** 2
    den = np.sum(ya ** 2) / si.n - fo2
    for ii in range(nvar):
        C = si.B.copy()
        C[:, ii] = si.A[:, ii]
        yc = pred(C)
        assert abs(res["first"][ii] - ((np.sum(ya * yc) / si.n - fo2) / den)) < 1e-7
        assert abs(res["total"][ii] - (1 - (np.sum(yb * yc) / si.n - fo2) / den)) < 1e-7
    ana = SobolIndices(nvar, problem="styblinski", ub=np.hstack([ub, ub]), lb=np.hstack([lb, lb])).analyze(True, True)
    assert np.max(np.abs(res["first"] - ana["first"])) < 0.08
    assert set(res["second"].keys()) == {"x1-x2", "x1-x3", "x2-x3"}


def _random_front(rng, k):
    f1 = np.sort(rng.uniform(0, 1, k))
    f2 = np.sort(rng.uniform(0, 1, k))[::-1]
    return np.column_stack([f1, f2])[rng.permutation(k)]


def test_exi2d_kernel_matches_reference_fixture_and_oracle():
    """SURVEY section 8(f) row 3 (EHVI part): the device cell sum against the reference's own exi2d
    outputs (tests/golden/ehvi2d.npz) and, on larger fronts and batches, against the oracle.
    Tolerance: rel 1e-9 of the largest improvement in the batch (cells are sums of clipped differences,
    so a candidate with a vanishing improvement carries the absolute error of the big cells)."""
    from oracle import ehvi_oracle as eo
    from bayesianopttools_b200.optim_tools.ehvi.exi2d import exi2d
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ehvi2d.npz"))
    for c in range(4):
        P, r, mu, s, ref = (g["%s%d" % (nm, c)] for nm in ("P", "r", "mu", "s", "ref"))
        got = exi2d(P, r, mu, s)
        assert np.max(np.abs(got - ref)) < 1e-9 * np.max(ref)
        assert abs(exi2d(P, r, mu[3], s[3]) - ref[3]) < 1e-9 * np.max(ref)      # the reference's scalar call shape
    rng = np.random.default_rng(11)
    for k, m in ((1, 5), (17, 3000), (40, 1000), (120, 300)):                     # k = 120: table read from global memory
        P = _random_front(rng, k)
        r = np.array([1.2, 1.1])
        mu = rng.uniform(-0.3, 1.3, (m, 2))
        s = rng.uniform(1e-3, 0.6, (m, 2))
        got, want = exi2d(P, r, mu, s), eo.exi2d(P, r, mu, s)
        assert np.max(np.abs(got - want)) < 1e-9 * np.max(want), (k, m)
        assert np.all(got >= 0)


def test_ehvicalc_batch_matches_oracle_predictions_and_argmax():
    """`ehvicalc` on two fitted objective models: per-candidate values equal exi2d of the ORACLE's
    predictions (variance handed over as the spread, as the reference does), the scalar call shape
    returns the same number, and the device argmax is the candidate numpy picks."""
    from oracle import ehvi_oracle as eo
    from oracle import kriging_oracle as ko
    from bayesianopttools_b200.optim_tools.ehvi.EHVIcomputation import ehvicalc, ehvi_best
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    rng = np.random.default_rng(3)
    n, d = 60, 2
    lb, ub = np.zeros(d), np.ones(d)
    X = rng.uniform(0, 1, (n, d))
    # a ZDT-like pair of conflicting objectives
    y1 = X[:, 0] + 0.1 * np.sin(6 * X[:, 1])
    y2 = 1.0 - np.sqrt(np.clip(X[:, 0], 0, 1)) + 0.5 * (X[:, 1] - 0.3) ** 2
    objs = []
    for y in (y1, y2):
        K = initkriginfo("single")
        K.update(X=X, y=y.reshape(-1, 1), lb=lb, ub=ub, nrestart=3, kernel=["gaussian"], nugget=-6, TrendOrder=0,
                 optimizer="lbfgsb")
        o = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
        o.train(disp=False)
        objs.append(o)
    Y = np.column_stack([y1, y2])
    nd = np.array([not np.any(np.all(Y <= Y[i], axis=1) & np.any(Y < Y[i], axis=1)) for i in range(n)])
    ypar = Y[nd]
    mobo = {"refpoint": np.array([1.3, 1.3])}
    xc = rng.uniform(0, 1, (5000, d))
    hv = ehvicalc(xc, ypar, mobo, objs)
    mu = np.zeros((len(xc), 2))
    var = np.zeros((len(xc), 2))
    for j, o in enumerate(objs):
        KI = o.KrigInfo
        fit = ko.nll(np.asarray(KI["Theta"]).ravel(), KI["X_norm"], KI["y"], KI["F"], full=True)
        p = ko.predict(ko.standardize_x(xc, lb, ub), KI["X_norm"], KI["y"], KI["F"], np.zeros((1, d), int), fit)
        mu[:, j], var[:, j] = p["pred"].ravel(), p["ssqr"].ravel()
    want = eo.exi2d(ypar, mobo["refpoint"], mu, var)
    assert np.max(want) > 0
    assert np.max(np.abs(-hv - want)) < 1e-8 * np.max(want)
    assert np.all((hv < 0) | (hv == np.finfo(float).tiny))          # exact zeros get the reference's penalty sign
    assert abs(ehvicalc(xc[17], ypar, mobo, objs) - hv[17]) <= 1e-12 * abs(hv[17])
    idx, val = ehvi_best(xc, ypar, mobo, objs)
    assert idx == int(np.argmin(hv)) and abs(val + hv[idx]) <= 1e-14 * abs(val)


def _hv2d(front, ref):
    f = front[np.argsort(front[:, 0])]
    h, prev = 0.0, ref[1]
    for a, b in f:
        if b < prev:
            h += (ref[0] - a) * (prev - b)
            prev = b
    return h


def test_mobo_ehvi_loop_with_believer_and_sweep():
    """the multi-objective loop end to end on the reference's demo problem
    (demo/MOBOdemo.py: Schaffer, 2 variables, 20 Halton samples): EHVI maximised by the device sweep,
    Kriging-believer multi-update (refits at fixed theta), enrichment and retraining.  Checks the
    bookkeeping the reference returns, that a believer refit equals the ORACLE's likelihood 'all' state on
    the augmented design, and that the dominated hypervolume grows."""
    from oracle import kriging_oracle as ko
    from bayesianopttools_b200.misc.sampling.samplingplan import sampling
    from bayesianopttools_b200.optim_tools.MOBO import MOBO
    from bayesianopttools_b200.optim_tools.searchpareto import paretopoint
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    from bayesianopttools_b200.testcase.analyticalfcn.cases import evaluate
    nvar, n = 2, 20
    lb, ub = -np.ones(nvar), np.ones(nvar)
    _, X = sampling("halton", nvar, n, result="real", upbound=ub, lobound=lb)
    Y = evaluate(X, "schaffer")
    objs = []
    for j in range(2):
        K = initkriginfo("single")
        K.update(X=X, y=Y[:, j].reshape(-1, 1), problem="schaffer", nrestart=3, ub=ub, lb=lb, optimizer="lbfgsb",
                 kernel=["gaussian"], nugget=-6, TrendOrder=0)
        o = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
        o.train(disp=False)
        objs.append(o)
    ref = np.array([1.5, 1.5])
    hv0 = _hv2d(paretopoint(Y)[0], ref)
    info = {"nup": 2, "nrestart": 2, "acquifunc": "ehvi", "acquifuncopt": "sweep", "ncandidates": 20000,
            "refpoint": ref, "filename": "/tmp/krigb200_mobo_test.mat"}       # candidates: device Sobol design
    opt = MOBO(info, objs, autoupdate=True, multiupdate=2, savedata=True)
    # one believer round on its own first: the refit state must equal the oracle on the augmented design
    opt.nup, opt.Xall, opt.yall = 0, X, Y.copy()
    opt.ypar, _ = paretopoint(opt.yall)
    xs, ys, ss, ms = opt.simultpredehvi()
    assert xs.shape == (2, nvar) and ys.shape == (2, 2) and ss.shape == (2, 2) and np.all(ms < 0)
    assert objs[0].KrigInfo["X"].shape[0] == n                       # the believers are copies
    # full loop
    xup, yup, sup, metric = opt.run(disp=False)
    assert xup.shape == (4, nvar) and yup.shape == (4, 2) and sup.shape == (4, 2) and len(metric) == 4
    assert np.max(np.abs(yup - evaluate(xup, "schaffer"))) < 1e-14    # enrichment stores true evaluations
    assert objs[0].KrigInfo["X"].shape[0] == n + 4 and objs[1].KrigInfo["y"].shape == (n + 4, 1)
    assert np.all(xup >= lb - 1e-12) and np.all(xup <= ub + 1e-12)
    hv1 = _hv2d(opt.ypar, ref)
    assert hv1 > hv0
    # retrained models reproduce the oracle's likelihood at their own hyper-parameters
    for o in objs:
        KI = o.KrigInfo
        fit = ko.nll(np.asarray(KI["Theta"]).ravel(), KI["X_norm"], KI["y"], KI["F"], full=True)
        assert abs(fit["NegLnLike"] - float(np.ravel(KI["NegLnLike"])[0])) <= 1e-9 * abs(fit["NegLnLike"])


def test_mobo_parego_loop_runs():
    """ParEGO branch of the multi-objective loop (MOBO.py:193-240): scalarised model + single-objective
    acquisition sweep, one and several (different weights) designs per iteration."""
    from bayesianopttools_b200.misc.sampling.samplingplan import sampling
    from bayesianopttools_b200.optim_tools.MOBO import MOBO
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    from bayesianopttools_b200.testcase.analyticalfcn.cases import evaluate
    nvar, n = 2, 16
    lb, ub = -np.ones(nvar), np.ones(nvar)
    _, X = sampling("halton", nvar, n, result="real", upbound=ub, lobound=lb)
    Y = evaluate(X, "schaffer")
    for multi in (0, 2):
        objs = []
        for j in range(2):
            K = initkriginfo("single")
            K.update(X=X, y=Y[:, j].reshape(-1, 1), problem="schaffer", nrestart=2, ub=ub, lb=lb,
                     optimizer="lbfgsb", kernel=["gaussian"], nugget=-6, TrendOrder=0)
            o = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
            o.train(disp=False)
            objs.append(o)
        info = {"nup": 2, "nrestart": 2, "acquifunc": "parego", "acquifuncopt": "sweep", "ncandidates": 5000,
                "rng": np.random.default_rng(1)}
        opt = MOBO(info, objs, autoupdate=True, multiupdate=multi, savedata=False)
        xup, yup, sup, metric = opt.run(disp=False)
        nnew = 2 * max(1, multi)
        assert xup.shape == (nnew, nvar) and yup.shape == (nnew, 2) and sup is None
        assert objs[0].KrigInfo["X"].shape[0] == n + nnew
        assert np.max(np.abs(yup - evaluate(xup, "schaffer"))) < 1e-14


def test_train_with_device_cmaes_reaches_the_multistart_optimum():
    """`Kriging.train` with optimizer='cmaes' runs the generation loop on the device (kb_train_cmaes) and
    ends at a likelihood as good as the multi-start L-BFGS-B search and the host-loop CMA-ES."""
    from bayesianopttools_b200.surrogate_models.kriging_model import Kriging
    from bayesianopttools_b200.surrogate_models.supports.initinfo import initkriginfo
    rng = np.random.default_rng(8)
    n, d = 80, 3
    X = rng.uniform(-5, 5, (n, d))
    y = 0.5 * np.sum(X ** 4 - 16 * X ** 2 + 5 * X, axis=1).reshape(-1, 1)
    got = {}
    for name, extra in (("lbfgsb", dict(optimizer="lbfgsb", nrestart=6)),
                        ("cmaes_device", dict(optimizer="cmaes", nrestart=2, seed=4)),
                        ("cmaes_host", dict(optimizer="cmaes", nrestart=2, seed=4, cmaes_device=False))):
        K = initkriginfo("single")
        K.update(X=X, y=y, lb=-5 * np.ones(d), ub=5 * np.ones(d), kernel=["gaussian"], nugget=-6, TrendOrder=0, **extra)
        o = Kriging(K, standardization=True, standtype="default", normy=False, trainvar=False)
        o.train(disp=False)
        got[name] = float(np.ravel(o.KrigInfo["NegLnLike"])[0])
    assert got["cmaes_device"] <= got["lbfgsb"] + 1e-4 * abs(got["lbfgsb"])
    assert abs(got["cmaes_device"] - got["cmaes_host"]) <= 1e-4 * abs(got["cmaes_host"])
