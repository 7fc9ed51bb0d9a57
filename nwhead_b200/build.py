"""Build libnw_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m nwhead_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnw_sm100.so")
SOURCES = ["nw_bank.cu", "nw_forward_sm100.cu", "nw_direct.cu", "nw_aux.cu"]
HEADERS = [os.path.join(CSRC, "nw_common.cuh"), os.path.join(os.path.dirname(HERE), "include", "nw_sm100.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libnw_sm100.so cannot be built")
    return nvcc


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB) -> str:
    """defines / out: developer A/B builds (e.g. defines=["NW_EPI_RSQRT=0"], out=".../libnw_sm100_sqrt.so", run with
    NW_B200_LIB pointing at it on the same GPU box)."""
    if not force and not defines and out == LIB and not is_stale():
        return LIB
    nvcc = find_nvcc()
    objdir = os.path.join(HERE, "build" if out == LIB else "build_" + os.path.basename(out))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
            "-cudart", "static", "-o", out, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libnw_sm100.so failed")
    return out


if __name__ == "__main__":
    defs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--define=")]
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv, defines=defs, out=outs[0] if outs else LIB)
    print(path)
