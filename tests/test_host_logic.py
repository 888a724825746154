"""CPU: host-side logic of the drop-in layer (no GPU compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, relerr


def test_library_loads_and_exports_every_declared_symbol():
    from bayesianopttools_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "krigb200.h")).read()
    declared = set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    assert lib.kb_version() >= 100
    assert isinstance(lib.kb_device_count(), int)


def test_header_enums_match_the_binding():
    """The enum values of include/krigb200.h the Python binding restates (normalisation, y map, refine mode)."""
    from bayesianopttools_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "krigb200.h")).read()
    enums = {k: int(v) for k, v in re.findall(r"\b(KB_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    assert (enums["KB_REFINE_AUTO"], enums["KB_REFINE_OFF"], enums["KB_REFINE_ON"]) == \
        (_lib.REFINE_AUTO, _lib.REFINE_OFF, _lib.REFINE_ON)
    for name, val in (("KB_NORM_DEFAULT", _lib.NORM_DEFAULT), ("KB_NORM_STD", _lib.NORM_STD), ("KB_NORM_NONE", _lib.NORM_NONE)):
        if name in enums:
            assert enums[name] == val, name


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayesianopttools_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_no_gpu_fails_loudly():
    from bayesianopttools_b200 import _lib
    if _lib.load().kb_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        _lib.DeviceModel(np.zeros((4, 2)), np.zeros(4), np.ones((4, 1)), ["gaussian"])
    with pytest.raises(RuntimeError):
        _lib.corr_matrix(np.zeros((2, 2)), np.zeros((2, 2)), np.ones(2))


def test_samplers_match_reference():
    from bayesianopttools_b200.misc.sampling.samplingplan import sampling
    g = load_golden("sampling")
    for nv, ns in ((2, 5), (10, 7), (3, 16)):
        assert np.array_equal(sampling("sobol", nv, ns)[0], g["sobol_%d_%d" % (nv, ns)])
        assert np.allclose(sampling("halton", nv, ns)[0], g["halton_%d_%d" % (nv, ns)], atol=1e-15)
    g = load_golden("krigdemo_nb")
    _, xh = sampling("sobol", 2, 5, result="real", upbound=np.array([5., 5.]), lobound=np.array([-5., -5.]))
    assert np.array_equal(xh, g["xhyp"])


def test_trend_and_standardize_match_reference():
    from bayesianopttools_b200.misc.sampling.samplingplan import standardize
    from bayesianopttools_b200.surrogate_models.supports.trendfunction import (compute_regression_mat,
                                                                                 polytruncation)
    g = load_golden("trend")
    for key in [k for k in g if k.startswith("idx_")]:
        _, o, v = key.split("_")
        idx = polytruncation(int(o), int(v), 1)
        assert np.array_equal(idx, g[key])
        nv = int(v)
        bound = np.vstack((-np.ones((1, nv)), np.ones((1, nv))))
        assert relerr(compute_regression_mat(idx, g["Xs_%s_%s" % (o, v)], bound, np.ones(nv)),
                      g["F_%s_%s" % (o, v)]) < 1e-13
    g = load_golden("ok_gaussian")
    Xn = standardize(g["X"], g["y"], type="default", range=np.vstack((g["lb"], g["ub"])))
    assert np.array_equal(Xn, g["X_norm"])


def test_kriginfocheck_hyperparameter_box():
    from bayesianopttools_b200.surrogate_models.kriging_model import kriginfocheck
    K = dict(type="kriging", nugget=[-8, -3], kernel=["gaussian", "matern32"], nrestart=3, optimizer="cmaes")
    K, scaling = kriginfocheck(K, -5, 5, 4)
    assert len(K["lbhyp"]) == 4 + 1 + 2 and K["nkernel"] == 2
    assert K["lbhyp"][4] == -8 and K["ubhyp"][4] == -3 and np.all(K["ubhyp"][5:] == 1)
    assert np.isclose(scaling[4], 0.5) and np.allclose(scaling[5:], 0.1)
    with pytest.raises(ValueError):
        kriginfocheck(dict(type="kriging", nugget=-6, optimizer="adam"), -5, 5, 2)


def test_cmaes_and_ga_minimise_with_batched_objective():
    from bayesianopttools_b200.optim_tools.cmaes import fmin_batched
    from bayesianopttools_b200.optim_tools.ga.uncGA import uncGA
    calls = []

    def sphere(P):
        calls.append(len(P))
        return np.sum((P - 0.7) ** 2, axis=1)

    bx, bf, es = fmin_batched(sphere, np.zeros(5), 2.0, bounds=[-5 * np.ones(5), 5 * np.ones(5)], popsize=16,
                              seed=1)
    assert bf < 1e-12 and np.allclose(bx, 0.7, atol=1e-5) and set(calls) == {16}
    bx, bf, _ = uncGA(sphere, lb=-5 * np.ones(3), ub=5 * np.ones(3), npop=60, maxg=80, seed=2)
    assert bf < 1e-2


def test_pls_rotations_match_reference_fixture():
    from bayesianopttools_b200.surrogate_models.kpls_model import pls1_rotations
    g = load_golden("kpls")
    W = pls1_rotations(g["X_norm"], g["y"], int(g["n_princomp"]))
    assert relerr(W ** 2, g["plscoeff"] ** 2) < 1e-9


def test_sobolnew_sampler_and_sobol_indices_match_reference(golden):
    """the Joe-Kuo `sobolnew` sampler (misc/sampling/sobol_new.py) and the
    SobolIndices estimators (sensitivity_analysis/sobol_ind.py) against outputs of the unmodified
    reference on the analytical Styblinski-Tang case (fixture written by /tmp-run of the reference,
    see oracle/make_golden.py::sobol_indices)."""
    from bayesianopttools_b200.sensitivity_analysis.sobol_ind import SobolIndices
    g = golden("sobol_indices")
    si = SobolIndices(3, krigobj=None, problem="styblinski", ub=g["ub"], lb=g["lb"])
    assert np.array_equal(si.A[:64], g["A_head"]) and np.array_equal(si.B[:64], g["B_head"])
    assert np.array_equal(si.A[-64:], g["A_tail"]) and np.array_equal(si.B[-64:], g["B_tail"])
    res = si.analyze(True, True, True)
    assert np.allclose(res["first"], g["first"], rtol=1e-12, atol=1e-14)
    assert np.allclose(res["total"], g["total"], rtol=1e-12, atol=1e-14)
    assert list(res["second"].keys()) == [str(k) for k in g["second_keys"]]
    assert np.allclose(list(res["second"].values()), g["second_vals"], rtol=1e-9, atol=1e-13)


def test_mobo_host_pieces_match_the_reference_fixture():
    """paretopoint (incl. duplicated rows, ties, 3 objectives, a single row) and the two-objective test
    problems against outputs of the unmodified reference (tests/golden/mobo_host.npz)."""
    from bayesianopttools_b200.optim_tools.searchpareto import paretopoint
    from bayesianopttools_b200.testcase.analyticalfcn.cases import evaluate
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mobo_host.npz"))
    for c in range(5):
        A, idx = paretopoint(g["B%d" % c])
        assert np.array_equal(np.asarray(A), g["A%d" % c]) and np.array_equal(np.asarray(idx), g["I%d" % c])
    for t in ("schaffer", "fonseca", "schaffer1"):
        assert np.max(np.abs(evaluate(g["X"], t) - g["y_" + t])) < 1e-15
        y1 = evaluate(g["X"][0], t)
        assert y1.shape == g["y1_" + t].shape and np.max(np.abs(y1 - g["y1_" + t])) < 1e-15


def test_paregopre_and_moboinfocheck_defaults():
    from bayesianopttools_b200.optim_tools.MOBO import moboinfocheck
    from bayesianopttools_b200.optim_tools.parego import paregopre
    Y = np.array([[0.0, 4.0], [1.0, 2.0], [2.0, 0.0]])
    # weights (0.3, 0.7) on objectives scaled to [0, 1]: max(w f) + 0.05 sum(w f)   (parego.py:14-29)
    F = np.array([[0.0, 1.0], [0.5, 0.5], [1.0, 0.0]])
    want = np.max(F * [0.3, 0.7], axis=1) + 0.05 * np.sum(F * [0.3, 0.7], axis=1)
    assert np.allclose(paregopre(Y, 3).ravel(), want, rtol=0, atol=1e-15)
    assert paregopre(Y).shape == (3, 1)
    info = moboinfocheck({"nup": 3}, True)
    assert info["acquifunc"] == "EHVI" and info["refpointtype"] == "dynamic" and info["acquifuncopt"] == "lbfgsb"
    assert info["nrestart"] == 1 and info["filename"] == "temporarydata.mat"
    assert moboinfocheck({"nup": 3, "refpoint": np.array([1.0, 1.0])}, True)["refpointtype"] == "static"
    assert moboinfocheck({"nup": 5}, False)["nup"] == 1
    with pytest.raises(ValueError):
        moboinfocheck({}, True)
    with pytest.raises(ValueError):
        moboinfocheck({"nup": 1, "acquifunc": "nsga"}, True)
    p = moboinfocheck({"nup": 1, "acquifunc": "parego"}, True)
    assert p["paregoacquifunc"] == "EI" and p["krignum"] == 1


def test_philox_restatement_known_answers():
    """oracle/philox.py against the known-answer vectors of the Random123 distribution (Salmon et al. 2011,
    kat_vectors: philox4x32-10) -- the generator the device population / CMA-ES sampling is checked against."""
    from oracle import philox
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want
    z = philox.population(np.arange(20000), 2, dist="normal", seed=11)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02 and np.all(np.isfinite(z))


def test_cmaes_restatement_minimises_known_functions():
    """oracle/cmaes_oracle.py (the checker of the device CMA-ES loop) finds the minimum of a shifted
    ellipsoid inside a box and respects the bounds when the minimum lies outside."""
    from oracle import cmaes_oracle
    c = np.array([0.3, -1.2, 2.0, 0.7])
    w = np.array([1.0, 10.0, 100.0, 3.0])
    f = lambda X: np.sum(w * (X - c) ** 2, axis=1)
    r = cmaes_oracle.run(f, np.zeros(4), 1.5, lb=-5 * np.ones(4), ub=5 * np.ones(4), seed=2)
    assert r["stop"] in (2, 3) and r["best_f"] < 1e-12 and np.max(np.abs(r["best_x"] - c)) < 1e-5
    r = cmaes_oracle.run(f, np.zeros(4), 1.5, lb=-np.ones(4), ub=np.ones(4), seed=2)
    assert np.all(np.abs(r["best_x"]) <= 1.0)
    assert np.max(np.abs(r["best_x"] - np.clip(c, -1, 1))) < 1e-4
    # a non-finite objective (what a non-positive-definite candidate returns) ranks last and is never the incumbent
    g = lambda X: np.where(X[:, 0] > 0.5, np.inf, np.sum((X + 0.2) ** 2, axis=1))
    r = cmaes_oracle.run(g, np.zeros(3), 1.0, seed=4, maxiter=200)
    assert np.isfinite(r["best_f"]) and r["best_x"][0] <= 0.5 and r["best_f"] < 1e-8


def test_lockstep_lbfgsb_driver_equals_scipy_restart_by_restart():
    """`lbfgsb_lockstep` drives scipy's reverse-communication routine for all restarts on one thread and evaluates
    their stencils as one population per round (kriging_model.py:358-368, :430 of the reference run the restarts one
    after another): every restart must end at scipy.optimize.minimize's bit-identical (x*, f*), including a start
    point outside the box and one that converges to a different basin."""
    from scipy.optimize import minimize
    from bayesianopttools_b200.surrogate_models.kriging_model import lbfgsb_lockstep
    rng = np.random.default_rng(0)
    n = 6
    lb, ub = -3 * np.ones(n), 2 * np.ones(n)
    A = rng.normal(size=(n, n))
    A = A @ A.T + np.eye(n)

    def batch(P):
        return np.array([0.5 * p @ A @ p + np.sum(np.sin(3 * p)) + 0.1 * np.sum(p ** 4) for p in P])

    def fg(x, eps=1e-3):
        h = np.full(n, eps)
        flip = (x + h > ub) & (x - h >= lb)
        h[flip] = -eps
        pts = np.vstack([x[None, :], x[None, :] + np.diag(h)])
        f = batch(pts)
        dx = (pts[1:, :] - x[None, :])[np.arange(n), np.arange(n)]
        return float(f[0]), (f[1:] - f[0]) / dx

    x0s = rng.uniform(-3, 2, (7, n))
    x0s[3] = ub + 0.5
    bx, bf, calls = lbfgsb_lockstep(batch, x0s, lb, ub)
    nfev = []
    for r in range(len(x0s)):
        res = minimize(fg, x0s[r], method="L-BFGS-B", jac=True, bounds=np.column_stack((lb, ub)))
        assert np.array_equal(res.x, bx[r]) and res.fun == bf[r], r
        nfev.append(res.nfev)
    assert calls == max(nfev)          # one population per round: as many device calls as the longest restart needs
