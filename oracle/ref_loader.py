"""Import the UNMODIFIED reference (fazaghifari/BayesianOptTools) in-process.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/make_golden.py`` (fixture
generation, in the build container where ``/root/reference`` is mounted) and by
nothing that ships.  ``/root/reference`` does not exist on the GPU box, so no
``-m gpu`` test, ``smoke()`` or ``bench.py`` may call :func:`load_reference`.

Recipe: the reference is not a package, imports four
third-party modules that are absent here (``cma``, ``matplotlib``,
``comet_ml``, ``anastruct``) and one private sklearn path that was removed
(``sklearn.cross_decomposition.pls_``), and opens a csv by a path relative to a
first-level sub-directory.  We stub the absent modules, alias the sklearn path
and chdir into ``demo/`` before importing.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("KADAL_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "surrogate_models"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def load_reference():
    """Returns a namespace with the reference's hot-path callables."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # --- stubs for absent third-party modules -------------------------------
    _stub("cma")
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib", cm=types.SimpleNamespace())
        _stub("matplotlib.pyplot")
        _stub("matplotlib.cm")
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        _stub("mpl_toolkits")
        _stub("mpl_toolkits.mplot3d", Axes3D=object)
    _stub("comet_ml", Experiment=object)
    _stub("anastruct", SystemElements=object)
    # --- sklearn private path alias (kpls_model.py:4) ------------------------
    import sklearn.cross_decomposition as _cd
    _stub("sklearn.cross_decomposition.pls_", PLSRegression=_cd.PLSRegression)
    cwd = os.getcwd()
    os.chdir(os.path.join(REFERENCE_ROOT, "demo"))
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from surrogate_models.kriging_model import Kriging
            from surrogate_models.kpls_model import KPLS
            from surrogate_models.supports.likelihood_func import likelihood
            from surrogate_models.supports.kernelfunc import calckernel
            from surrogate_models.supports.prediction import prediction
            from surrogate_models.supports.trendfunction import (
                polytruncation, compute_regression_mat)
            from surrogate_models.supports.initinfo import initkriginfo
            from surrogate_models.supports.krigloocv2 import loocv2
            from misc.sampling.samplingplan import sampling, standardize, realval
            from testcase.analyticalfcn.cases import evaluate as evaluate_analytic
            from reliability_analysis.akmcs import AKMCS, mcpopgen
            from testcase.RA.testcase import evaluate as evaluate_ra
    finally:
        os.chdir(cwd)
    return types.SimpleNamespace(**{k: v for k, v in locals().items()
                                    if not k.startswith("_") and k not in ("cwd", "warnings")})
