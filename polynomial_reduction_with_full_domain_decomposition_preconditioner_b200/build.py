"""
Builds libprfdd_b200.so (CUDA kernels for sm_100a + C++ host classes + C ABI) and the `poisson`
driver, in-tree, with nvcc.  No torch involved: the library is plain CUDA runtime + (dlopen'd) NCCL.

    python -m polynomial_reduction_with_full_domain_decomposition_preconditioner_b200.build
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libprfdd_b200.so")
EXE = os.path.join(HERE, "poisson")
COMPAT = os.path.join(HERE, "libprfdd_compat.so")   # the reference's own five extern "C" names (include/prfdd_compat.h)

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--extended-lambda", "-I/usr/include"]

CU_SOURCES = ["k_vector.cu", "k_sem.cu", "k_sparse.cu", "k_krylov.cu"]
CPP_SOURCES = ["host/special_functions.cpp", "host/globals.cpp", "host/comm.cpp", "host/mesh.cpp", "host/capi.cpp"]


def _newer(src, obj, extra_deps):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + extra_deps)


def _headers():
    out = [os.path.join(HERE, "..", "include", "prfdd_b200.h")]
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".hpp", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    return out


def _compile(src):
    obj = os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")
    full = os.path.join(CSRC, src)
    if _newer(full, obj, _headers()):
        cmd = ["nvcc"] + ARCH + COMMON + ["-x", "cu", "-c", full, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build(verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = CU_SOURCES + CPP_SOURCES
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_compile, srcs))
    if (not os.path.exists(LIB)) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = ["nvcc"] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    main = os.path.join(CSRC, "host", "poisson.cpp")
    if os.path.exists(main) and ((not os.path.exists(EXE)) or os.path.getmtime(main) > os.path.getmtime(EXE) or os.path.getmtime(LIB) > os.path.getmtime(EXE)):
        cmd = ["nvcc"] + ARCH + ["-O2", "-std=c++17", main, "-o", EXE, "-L" + HERE, "-lprfdd_b200", "-Xlinker", "-rpath=$ORIGIN", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("poisson link failed:\n%s\n%s" % (r.stdout, r.stderr))
    csrc = os.path.join(CSRC, "compat.cpp")
    if (not os.path.exists(COMPAT)) or os.path.getmtime(csrc) > os.path.getmtime(COMPAT) or os.path.getmtime(LIB) > os.path.getmtime(COMPAT):
        cmd = ["nvcc"] + ARCH + ["-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared", "-x", "cu", csrc, "-o", COMPAT,
               "-L" + HERE, "-lprfdd_b200", "-Xlinker", "-rpath=$ORIGIN", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("compat link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(verbose=True)
