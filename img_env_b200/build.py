"""Builds libimgenv_b200.so (hand-written CUDA for sm_100a) in-tree with nvcc.

No JIT cache and no torch extension machinery: one explicit nvcc command so the built .so
travels with the source tree and `ncu`/`cuobjdump` see exactly what runs."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libimgenv_b200.so")
SOURCES = ["imgenv.cu"]
HEADERS = ["state.cuh", "tfmath.cuh", "kin.cuh", "view.cuh", "dyn.cuh", "orca.cuh", "sfmtree.cuh", "host_tables.h", "sampler.h"]


def nvcc_path():
    for p in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(HERE, "..", "include", "imgenv.h")]
    return any(os.path.getmtime(f) > t for f in deps)


def build_variant(out, defs):
    """Experiment helper: the same sources with extra -D macros into another .so (selected with IMGENV_LIB_PATH)."""
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", out] + ["-D" + d for d in defs] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    return out


def build(force=False, verbose=False):
    if os.environ.get("IMGENV_LIB_PATH"):
        return os.environ["IMGENV_LIB_PATH"]
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
