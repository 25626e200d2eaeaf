"""Builds the CUDA libraries in-tree with nvcc for sm_100a (no JIT cache: the .so files travel with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build(force=False, verbose_ptxas=False):
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".hpp", ".h"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    extra = ["-Xptxas", "-v"] if verbose_ptxas else []
    extra += os.environ.get("AIRICE_EXTRA_NVCC", "").split()       # development: -D knobs of the kernels (tools/*_probe)

    core = os.path.join(LIBDIR, "libairice_b200.so")
    core_src = [os.path.join(CSRC, f) for f in ("kernels.cu", "capi.cu", "atmosphere.cc")]
    # translation units that must round like the reference's x86 build (no FMA contraction): the in-ice solver replays
    # the reference's iterations; the old-table lookup is plain arithmetic on stored doubles
    nofma = [("inice_kernels.cu", "inice_kernels.o"), ("oldtable_kernels.cu", "oldtable_kernels.o")]
    nofma_src = [os.path.join(CSRC, a) for a, _ in nofma]
    nofma_obj = [os.path.join(LIBDIR, b) for _, b in nofma]
    if force or _newer(core, core_src + nofma_src + headers):
        for src, obj in zip(nofma_src, nofma_obj):
            _run(["nvcc"] + NVCC_FLAGS + extra + ["-fmad=false", "-c", "-o", obj, src])
        _run(["nvcc"] + NVCC_FLAGS + extra + ["-shared", "-o", core] + core_src + nofma_obj)

    # source-compatible C++ API (namespace MultiRayAirIceRefraction) on top of the C ABI
    compat = os.path.join(LIBDIR, "libMultiRayAirIceRefraction.so")
    compat_src = [os.path.join(CSRC, "multiray_compat.cc")]
    if os.path.exists(compat_src[0]) and (force or _newer(compat, compat_src + headers + [core])):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", compat] +
             compat_src + ["-L" + LIBDIR, "-lairice_b200", "-Wl,-rpath,$ORIGIN"])

    # python-wrapper C ABI: libAirIceRayTracing.so exporting Py_TraceIceToAir, next to AirIceRayTracing.py
    pw_dir = os.path.join(HERE, "pythonwrapper")
    pw = os.path.join(pw_dir, "libAirIceRayTracing.so")
    pw_src = [os.path.join(CSRC, "pywrap.cc")]
    if os.path.exists(pw_src[0]) and (force or _newer(pw, pw_src + headers + [core])):
        os.makedirs(pw_dir, exist_ok=True)
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", pw] + pw_src +
             ["-L" + LIBDIR, "-lairice_b200", "-Wl,-rpath,$ORIGIN/../lib"])
    return core


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
