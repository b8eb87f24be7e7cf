"""
ORACLE tooling (test infrastructure, NOT product code): ctypes access to oracle/liboracle.so
(speclib.c + kernels.c) and, when present, to the translated reference kernels in oracle/_ref/.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref = {}

f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    """Compile liboracle.so, liboracle_omp.so (and oracle/_ref when /root/reference is mounted)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so", "liboracle_omp.so", "liboracle_f32.so"])
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"], stdout=subprocess.DEVNULL)


_threads = 1


def set_threads(n):
    """n > 1: the @outer loops of every kernel run on n host threads (liboracle_omp.so: the same source compiled with OpenMP;
    the OCCA OpenMP analogue of config.hpp:34-36); bit-identical results.  Call before the first kernel call."""
    global _lib, _threads
    n = max(1, int(n))
    if (n > 1) != (_threads > 1):
        _lib = None
    _threads = n
    if n > 1:
        os.environ["OMP_NUM_THREADS"] = str(n)
        lib()


def lib():
    global _lib
    if _lib is None:
        name = "liboracle_omp.so" if _threads > 1 else "liboracle.so"
        path = os.path.join(HERE, name)
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.oracle_hgll.restype = C.c_double
        _lib.o_serial_sum.restype = C.c_double
    return _lib


_lib32 = None


def lib32():
    """the kernels compiled with DType = float (`Float float`, AMG/config.hpp:4): only the AMG loops are used from it"""
    global _lib32
    if _lib32 is None:
        path = os.path.join(HERE, "liboracle_f32.so")
        if not os.path.exists(path):
            build()
        _lib32 = C.CDLL(path)
    return _lib32


def ref(dim):
    """The reference's own OKL kernels compiled for CPU (None if oracle/_ref was never built)."""
    if dim not in _ref:
        path = os.path.join(HERE, "_ref", "libref_okl_%dd.so" % dim)
        _ref[dim] = C.CDLL(path) if os.path.exists(path) else None
    return _ref[dim]


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def ptr_table(arrays):
    """device-pointer-table analogue (domain.tpp:65-67, 221-224): array of double*."""
    t = (C.c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        t[i] = a.ctypes.data
    return t


# ---- speclib -------------------------------------------------------------------------------
def zwgll(n):
    z = np.zeros(n); w = np.zeros(n)
    lib().oracle_zwgll(ptr(z), ptr(w), C.byref(C.c_int(n)))
    return z, w


def dgll(z, n):
    """Returns D with D[i, j] = dl_j/dxi(xi_i), exactly as domain.tpp:312-314 reads it.  z may be modified."""
    D = np.zeros(n * n); Dt = np.zeros(n * n)
    # reference call: dgll_(Dt_gll, D_gll, r, &n, &n)
    lib().oracle_dgll(ptr(Dt), ptr(D), ptr(z), C.byref(C.c_int(n)), C.byref(C.c_int(n)))
    return D.reshape(n, n)


def hgll(j, zval, zgll, n):
    zz = C.c_double(zval)
    return lib().oracle_hgll(C.byref(C.c_int(j)), C.byref(zz), ptr(zgll), C.byref(C.c_int(n)))
