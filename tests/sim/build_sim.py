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
    srcs = [os.path.join(HERE, f) for f in ("fake_cudart.cc", "sim_kernels.cc", "sim_rsp.cc", "hostemu_tree.cc", "hostemu_dense.cc")]
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-pthread", "-fno-strict-aliasing",
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


def build_sanitized(verbose=False):
    """The engine's host code AND the stand-in runtime recompiled with AddressSanitizer + UBSan
    (tests/sim/_build/asan/).  The device code in the objects is untouched (nothing executes it here); what is
    checked is the host side: replica / state bookkeeping, work-list construction, the C ABI's buffer handling.
    Run python with LD_PRELOAD=<libasan> (see `sanitizer_env`)."""
    sys.path.insert(0, PKG)
    import importlib.util
    spec = importlib.util.spec_from_file_location("mxkv_build", os.path.join(PKG, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    out = os.path.join(OUT, "asan")
    os.makedirs(out, exist_ok=True)
    san = ["-fsanitize=address", "-fsanitize=undefined", "-fno-sanitize=vptr", "-fno-omit-frame-pointer",
           "-fno-sanitize-recover=undefined"]
    fake = os.path.join(out, "libcudart.so.12")
    srcs = [os.path.join(HERE, f) for f in ("fake_cudart.cc", "sim_kernels.cc", "sim_rsp.cc", "hostemu_tree.cc", "hostemu_dense.cc")]
    deps = srcs + [os.path.join(PKG, "csrc", f) for f in os.listdir(os.path.join(PKG, "csrc"))] + \
           [os.path.join(HERE, "sim.h"), os.path.abspath(__file__)]
    sim = os.path.join(out, "libmxkv_b200_sim.so")
    if os.path.exists(sim) and os.path.exists(fake) and \
            all(os.path.getmtime(d) <= os.path.getmtime(sim) for d in deps):
        return sim, fake
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wno-unknown-pragmas", "-pthread",
           "-fno-strict-aliasing"] + \
          san + \
          ["-I", os.path.join(PKG, "csrc"), "-I", os.path.join(CUDA, "include"), "-I", HERE] + srcs + \
          ["-o", fake, "-Wl,-soname,libcudart.so.12", "-Wl,--version-script=" + os.path.join(HERE, "cudart.map")]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    procs, objs = [], []
    for src in b.SOURCES:
        obj = os.path.join(out, src.rsplit(".", 1)[0] + ".o")
        flags = [f for f in b.FLAGS if f != "-O3"]
        cmd = [b.NVCC] + flags + ["-O1", "-g", "-Xcompiler", ",".join(san), "-x", "cu", "-c", os.path.join(PKG, "csrc", src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("sanitized build failed on %s:\n%s" % (src, o.decode()))
    cmd = ["g++", "-shared", "-o", sim] + objs + ["-fsanitize=address,undefined", "-L", out, "-l:libcudart.so.12",
                                                  "-Wl,--disable-new-dtags", "-Wl,-rpath," + out]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return sim, fake


def sanitizer_env():
    """Environment additions for a python process that loads the sanitized libraries."""
    def lib(name):
        return subprocess.check_output(["gcc", "-print-file-name=" + name], text=True).strip()
    return {"LD_PRELOAD": lib("libasan.so") + ":" + lib("libubsan.so"),
            # python itself leaks by design; the interceptors must not trip over its allocator either
            "ASAN_OPTIONS": "detect_leaks=0:abort_on_error=1:allocator_may_return_null=1:handle_segv=0",
            "UBSAN_OPTIONS": "print_stacktrace=1:halt_on_error=1",
            "PYTHONMALLOC": "malloc"}


if __name__ == "__main__":
    if "--asan" in sys.argv:
        print(build_sanitized(verbose=True))
    else:
        print(build(verbose=True))
