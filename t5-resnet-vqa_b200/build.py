"""Build libvqa_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvqa_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "vqa_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug=False):
    """Compile every .cu under csrc/ into one shared library; returns its path.  debug=True builds the diagnostic
    variant libvqa_b200_dbg.so (-DVQA_GEMM_DEBUG: phase stamps and load-only / MMA-only modes of the GEMM kernel, used
    by tools/gemm_timing.py and tools/gemm_dbg.py through VQA_B200_LIB)."""
    if debug:
        return _build_into(os.path.join(HERE, "libvqa_b200_dbg.so"), os.path.join(HERE, "build_dbg"),
                           FLAGS + ["-DVQA_GEMM_DEBUG"], True, verbose)
    if not force and not _stale():
        return LIB
    return _build_into(LIB, os.path.join(HERE, "build"), FLAGS, force, verbose)


def _build_into(LIB, objdir, FLAGS, force, verbose):
    objs = []
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(os.path.join(CSRC, f))
                                                for f in os.listdir(CSRC) if f.endswith(".cuh") or f == os.path.basename(src))
                and os.path.getmtime(obj) > os.path.getmtime(os.path.join(HERE, "..", "include", "vqa_b200.h"))):
            continue
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed for %s" % src)
    cmd = [NVCC, "-shared", "-cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))
