"""In-tree build of libphylomap_b200.so (sm_100a only): one nvcc per translation unit, in parallel, then a link.

    python -m phylomap_b200.build [--force]
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libphylomap_b200.so")
UNITS = ["pm_host", "pm_sweep_f64x", "pm_sweep_f64x_gen", "pm_sweep_f64", "pm_sweep_f64_gen", "pm_sweep_f32",
         "pm_sweep_f32_gen"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp"))]
    out.append(os.path.join(HERE, "..", "include", "phylomap_b200.h"))
    return out


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _compile(unit, verbose):
    src = os.path.join(CSRC, unit + ".cu")
    obj = os.path.join(OBJ, unit + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, r.stdout, r.stderr))
    return unit, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    deps = _deps()
    todo = [u for u in UNITS
            if force or _stale(os.path.join(OBJ, u + ".o"), [os.path.join(CSRC, u + ".cu")] + deps)]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for unit, log in ex.map(lambda u: _compile(u, verbose), todo):
                if verbose:
                    sys.stderr.write("== %s ==\n%s\n" % (unit, log))
    objs = [os.path.join(OBJ, u + ".o") for u in UNITS]
    if todo or _stale(LIB, objs):
        subprocess.check_call([_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
