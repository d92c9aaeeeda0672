"""In-tree build of the CUDA library (sm_100a only): python -m sqp_solver_b200.build

Produces sqp_solver_b200/libsqp_b200.so from csrc/*.cu with nvcc. The .so is git-ignored but
travels to the GPU box with the repo snapshot. nvcc cross-compiles here without a GPU."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsqp_b200.so")
TOOL = os.path.join(HERE, "build", "batch_sqp_bench")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-Wall", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets", "-ccbin", "/usr/bin/g++"]


# per-file flags: the thread-per-QP kernel restates the CPU path's arithmetic operation by operation, so no FMA contraction there
PER_FILE = {"qp_small.cu": ["-fmad=false"]}


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    if not os.path.exists(TOOL):
        return True
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h")) + \
        glob.glob(os.path.join(HERE, "host", "*", "*.*pp")) + glob.glob(os.path.join(HERE, "host", "*", "*", "*.*pp"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    if not force and not stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        cmd = [NVCC] + FLAGS + PER_FILE.get(os.path.basename(src), []) + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            sys.stderr.write(out)
        if pr.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ccbin", "/usr/bin/g++", "-Wno-deprecated-gpu-targets"]
    subprocess.check_call(cmd)
    build_tools()
    return LIB


def build_tools():
    """Host-side tools on top of the library: the batched-SQP benchmark driver (BASELINE config 4; bench.py runs it)."""
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-std=c++11", "-O2", "-Wall", "-fopenmp", os.path.join(HERE, "host", "tools", "batch_sqp_bench.cpp"),
                           "-o", TOOL, "-L" + HERE, "-lsqp_b200", "-Wl,-rpath,$ORIGIN/..", "-L/usr/local/cuda/lib64",
                           "-Wl,-rpath,/usr/local/cuda/lib64"])
    return TOOL


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
