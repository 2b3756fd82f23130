"""Build the CUDA engine (sm_100a only) into hierarchical-irm_b200/libhirm_engine.so."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "engine.cu")
SRC_HOST = [os.path.join(HERE, "csrc", f) for f in ("ingest.cc", "relation_moves.cc", "sweep_rounds.cc", "ingest_gpu.cu")]
LIB = os.path.join(HERE, "libhirm_engine.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in
        ("engine.cu", "engine_types.cuh", "device_math.cuh", "generic_sweep.cuh", "dense_sweep.cuh", "pregroup.cuh", "aux_kernels.cuh", "ingest.cc", "relation_moves.cc", "sweep_rounds.cc", "ingest_gpu.cu")] + \
       [os.path.join(os.path.dirname(HERE), "include", "hirm_engine.h")]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    return None


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build the CUDA engine (there is no CPU fallback)")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177", "-o", LIB, SRC] + SRC_HOST
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
