"""Builds the simulated CUDA runtime (tests/sim/_build/libcudart.so.12: fake_cudart.cc + the kernel
emulators) and a variant of the engine linked against it (tests/sim/_build/libmxkv_b200_sim.so: the SAME
object files as the product library, only with the CUDA runtime linked dynamically so that the stand-in is
picked up).  Test infrastructure; see README.md in this directory."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "incubator-mxnet_b200")
OUT = os.path.join(HERE, "_build")
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")


def build(verbose=False):
    sys.path.insert(0, PKG)
    import importlib.util
    spec = importlib.util.spec_from_file_location("mxkv_build", os.path.join(PKG, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build()                                   # the product library and its object files
    os.makedirs(OUT, exist_ok=True)
    fake = os.path.join(OUT, "libcudart.so.12")
    srcs = [os.path.join(HERE, f) for f in ("fake_cudart.cc", "sim_kernels.cc", "sim_rsp.cc")]
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas",
           "-I", os.path.join(PKG, "csrc"), "-I", os.path.join(CUDA, "include"), "-I", HERE] + srcs + \
          ["-o", fake, "-Wl,-soname,libcudart.so.12", "-Wl,--version-script=" + os.path.join(HERE, "cudart.map")]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    objs = [os.path.join(PKG, "build", f) for f in sorted(os.listdir(os.path.join(PKG, "build"))) if f.endswith(".o")]
    sim = os.path.join(OUT, "libmxkv_b200_sim.so")
    cmd = [b.NVCC, "-shared", "-cudart", "shared", "-o", sim] + objs + \
          ["-gencode", "arch=compute_100a,code=sm_100a",
           # DT_RPATH (not RUNPATH): the stand-in must win over an LD_LIBRARY_PATH that holds the real runtime
           "-Xlinker", "--disable-new-dtags", "-Xlinker", "-rpath," + OUT]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return sim, fake


if __name__ == "__main__":
    print(build(verbose=True))
