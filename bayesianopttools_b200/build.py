"""Build libkrigb200.so (sm_100a only) in-tree with nvcc.

``python -m bayesianopttools_b200.build`` or ``build()`` from
``__graft_entry__``.  nvcc cross-compiles without a GPU; the resulting ``.so``
stays in the tree (git-ignored) so it travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libkrigb200.so")
SOURCES = ["kb_corr.cu", "kb_chol.cu", "kb_gls.cu", "kb_predict.cu", "kb_rng.cu", "kb_ehvi.cu", "kb_cmaes.cu", "kb_small.cu", "kb_peak.cu", "kb_pools.cu", "kb_api.cu", "kb_api_predict.cu", "kb_api_train.cu"]
HEADERS = ["kb_common.cuh", "kb_internal.h", "kb_diag.cuh", "kb_philox.cuh", "kb_api_internal.h", os.path.join("..", "..", "include", "krigb200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build the sm_100a kernels")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, extra_flags=()):
    """variant / extra_flags: development aid -- an A/B library `libkrigb200_<variant>.so` compiled with extra
    -D flags next to the product library (select it at run time with KRIGB200_LIB=<path>)."""
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build" if not variant else "build_" + variant)
    LIB = os.path.join(CSRC, "libkrigb200.so" if not variant else "libkrigb200_%s.so" % variant)
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        path = os.path.join(CSRC, src)
        if force or _stale(obj, [path] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + ["-c", path, "-o", obj]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed for %s:\n%s" % (src, res.stderr))
            with open(obj + ".ptxas.log", "w") as fh:
                fh.write(res.stderr)
            if verbose:
                sys.stderr.write(res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s" % res.stderr)
    return LIB


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var[0] if var else None,
                extra_flags=[a for a in sys.argv[1:] if a.startswith("-D")]))
