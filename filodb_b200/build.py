"""Builds filodb_b200/libfilo_b200.so (the C-ABI of include/filo_b200.h) in-tree with nvcc for sm_100a.

    python -m filodb_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false: the JVM never contracts a*b+c, parity with the reference is bit-exact
per series only without FMA contraction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# FILO_BUILD_OUT: build a variant (e.g. FILO_NVCC_EXTRA=-DFILO_HIST_PROF) into another file, with its own object directory
OUT = os.path.abspath(os.environ["FILO_BUILD_OUT"]) if os.environ.get("FILO_BUILD_OUT") else os.path.join(HERE, "libfilo_b200.so")
OBJ = os.path.join(HERE, "csrc", "_obj" if not os.environ.get("FILO_BUILD_OUT") else "_obj_variant")
SOURCES = ["scan_kernels.cu", "hist_kernels.cu", "hist_kernels2.cu", "synth_kernels.cu", "result_kernels.cu", "capi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-DFILO_BUILDING"] + os.environ.get("FILO_NVCC_EXTRA", "").split()


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp"))] + \
           [os.path.join(os.path.dirname(HERE), "include", "filo_b200.h"), os.path.abspath(__file__)]


def build(force=False, verbose=False):
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in _deps()):
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OBJ, s + ".o")
        objs.append(o)
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (s, out))
        if verbose:
            print(out)
    cmd = ["nvcc", "-shared", "-Wno-deprecated-gpu-targets", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    build_jni()
    return OUT


JNI_OUT = os.path.join(HERE, "libfilo_b200_jni.so")


def build_jni():
    """The JNI shim (csrc/jni_shim.cpp) over the C-ABI, against the minimal jni.h of csrc/jni_stub (no JDK in this image; with one:
    FILO_JNI_INCLUDE=$JAVA_HOME/include builds against the real header)."""
    if os.environ.get("FILO_BUILD_OUT"):
        return None
    inc = os.environ.get("FILO_JNI_INCLUDE")
    flags = ["-DFILO_USE_SYSTEM_JNI", "-I", inc, "-I", os.path.join(inc, "linux")] if inc else []
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wextra", "-Wno-unused-parameter"] + flags + \
          [os.path.join(CSRC, "jni_shim.cpp"), "-o", JNI_OUT, "-L", HERE, "-lfilo_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return JNI_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
