"""KrigInfo dict factory (reference: surrogate_models/supports/initinfo.py:49-139)."""


def init_kriginfo(objectives=1):
    if not isinstance(objectives, int):
        raise TypeError("The objectives parameter must be an integer.")
    if objectives < 1:
        raise ValueError("The number of objectives must be 1 or greater.")
    info = dict(X=None, nvar=None, problem="branin", nsamp=None, nrestart=5, kernel="gaussian", nugget=-6,
                standardization=True)
    info["krignum"] = None if objectives == 1 else 1
    info["multiobj"] = objectives > 1
    for key in ("y", "lb", "ub", "Theta", "U", "Psi", "BE", "y_mean", "y_std", "SigmaSqr", "idx", "F", "wgkf",
                "plscoeff", "NegLnLike"):
        info[key] = [0] * objectives
    return info


def initkriginfo(type="single", objective=1):
    if type.lower() == "single":
        return init_kriginfo(1)
    return init_kriginfo(objective)
