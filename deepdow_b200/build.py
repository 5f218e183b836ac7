"""Build libddw_b200.so in-tree with nvcc for sm_100a (no GPU needed: nvcc cross-compiles)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(HERE, "libddw_b200.so")
SOURCES = ["api.cu", "markowitz_fwd_sparse_f32.cu", "markowitz_fwd_sparse_f64.cu", "markowitz_bwd_sparse_f32.cu",
           "markowitz_bwd_sparse_f64.cu", "markowitz_fwd_f32.cu", "markowitz_fwd_f64.cu", "markowitz_bwd_f32.cu", "markowitz_bwd_f64.cu", "markowitz_gram_tc.cu", "markowitz_qwarp.cu",
           "rowproj.cu", "covariance.cu", "eig_block.cu", "risk_budgeting.cu", "host_pipeline.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-I", INCLUDE, "-I", CSRC]
if os.environ.get("DDW_PHASE_TIMING"):      # debug build: per-phase cycle counters in the sparse kernels
    NVCC_FLAGS.append("-DDDW_PHASE_TIMING")
if os.environ.get("DDW_GS_TIMING"):         # debug build: per-role cycle counters in the tensor-core Gram kernel
    NVCC_FLAGS.append("-DDDW_GS_TIMING")


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libddw_b200.so cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps(path, seen=None):
    """The source and every header it includes (recursively, quoted includes only)."""
    import re
    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for inc in re.findall(r'^\s*#\s*include\s+"([^"]+)"', open(path).read(), flags=re.M):
        for d in (os.path.dirname(path), CSRC, INCLUDE):
            cand = os.path.join(d, inc)
            if os.path.exists(cand):
                _deps(cand, seen)
                break
    return seen


def build(force=False, verbose=False):
    """Compile every .cu to an object (in parallel, only those whose own includes changed) and link the shared library."""
    # debug (phase-timing) objects live apart from the product objects, so the two flag sets never mix
    objdir = os.path.join(HERE, "build_dbg" if os.environ.get("DDW_PHASE_TIMING") else ("build_gst" if os.environ.get("DDW_GS_TIMING") else "build"))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, sorted(_deps(s))):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([nvcc] + NVCC_FLAGS + extra + ["-c", s, "-o", o])
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            results = list(ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs))
        for cmd, r in zip(jobs, results):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
    stamp = os.path.join(HERE, "build", ".linked_from")
    last = open(stamp).read() if os.path.exists(stamp) else ""
    if jobs or force or _stale(LIB, objs) or last != objdir:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
        os.makedirs(os.path.dirname(stamp), exist_ok=True)
        open(stamp, "w").write(objdir)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
