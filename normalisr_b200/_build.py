"""Build the CUDA library in-tree: normalisr_b200/libnsr_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnsr_b200.so")
SOURCES = ["api.cu", "basis.cu", "binnet.cu", "grouped.cu", "lcpm.cu", "normvar.cu", "sympinv.cu", "residual.cu", "contract_simt.cu", "contract_umma.cu"]
HEADERS = ["nsr_common.cuh", "epilogue.cuh", "pvalue.cuh", os.path.join("..", "..", "include", "normalisr_b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    cmd = [_nvcc(), "-t", "4", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v" if verbose else "-warn-spills",
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libnsr_b200.so:\n" + res.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
