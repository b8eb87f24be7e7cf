"""ctypes binding of include/prfdd_b200.h (libprfdd_b200.so).  Thin: no arithmetic happens here."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libprfdd_b200.so")
_lib = None

# ids of include/prfdd_b200.h
Q = dict(DIM=1, NUM_LOCAL_ELEMENTS=2, NUM_LOCAL_POINTS=3, NUM_LOCAL_NODES=4, NUM_BDARY_NODES=5, NUM_TOTAL_ELEMENTS=6,
         NUM_GLOBAL_NODES=7, SUB_NUM_POINTS=8, SUB_NUM_DOFS=9, SUB_NUM_EXTENDED_DOFS=10, SUP_NUM_DOFS=11,
         SUP_NUM_EXTENDED_DOFS=12, NUM_VALUES=13, NUM_DOFS=14, AMG_NUM_LEVELS=15, INNER_ITERATIONS=16,
         GPU_LAUNCHES_PER_PRECOND=17)
A = dict(NODE_OF_POINT=100, BOUNDARY_NODES=101, ASSEMBLED_WEIGHT=102, D_HAT=103, U=104, U_STAR=105, F=106,
         SUB_Q_PTR=107, SUB_Q_COL=108, SUB_Q_VAL=109, SUB_ELEMENT_IDS=110, SUB_ELEMENT_DEGREE=111, SUB_DOF_NUM=112,
         AMG_LEVEL_ROWS=113, AMG_LEVEL_NNZ=114, AMG_CHEBY_COEFS=115, A_FEM_PTR=116, A_FEM_COL=117, A_FEM_VAL=118,
         NORM_WEIGHT=119, INNER_WEIGHT=120)
A_DTYPE = dict(NODE_OF_POINT=np.int32, BOUNDARY_NODES=np.int64, SUB_Q_PTR=np.int32, SUB_Q_COL=np.int32,
               SUB_ELEMENT_IDS=np.int32, SUB_ELEMENT_DEGREE=np.int32, SUB_DOF_NUM=np.int64, AMG_LEVEL_ROWS=np.int32,
               AMG_LEVEL_NNZ=np.int32, A_FEM_PTR=np.int32, A_FEM_COL=np.int32)
APPLY = dict(STIFFNESS=200, DSSUM=201, DSSUM_WEIGHTED=202, PRECONDITIONER=203, SUB_STIFFNESS=204, LOW_ORDER=205,
             VCYCLE=206, TREE=207)


class Options(C.Structure):
    _fields_ = [("poly_degree", C.c_int), ("poly_reduction", C.c_int), ("subdomain_overlap", C.c_int),
                ("superdomain_overlap", C.c_int), ("use_preconditioner", C.c_int), ("preconditioner_type", C.c_int),
                ("inner_num_vectors", C.c_int), ("inner_max_iterations", C.c_int), ("num_vcycles", C.c_int),
                ("cheby_order", C.c_int), ("use_cuda_graph", C.c_int), ("proc_id", C.c_int), ("num_procs", C.c_int),
                ("nccl_unique_id", C.c_void_p), ("outer_tolerance", C.c_double), ("inner_tolerance", C.c_double),
                ("outer_max_iterations", C.c_int), ("outer_num_vectors", C.c_int), ("verbose", C.c_int), ("amg_coarsening", C.c_int), ("amg_precision", C.c_int), ("device_outer_loop", C.c_int)]


def lib():
    """Loads libprfdd_b200.so.  Raises (no fallback) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libprfdd_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.prfdd_version.restype = C.c_char_p
        L.prfdd_error_string.restype = C.c_char_p
        L.prfdd_launch_count.restype = C.c_longlong
        L.prfdd_solver_query.restype = C.c_longlong
        L.prfdd_solver_get_array.restype = C.c_longlong
        L.prfdd_solver_timer_total.restype = C.c_double
        L.prfdd_hgll.restype = C.c_double
        L.prfdd_solver_query.argtypes = [C.c_void_p, C.c_int]
        L.prfdd_solver_get_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
        L.prfdd_solver_timer_total.argtypes = [C.c_void_p, C.c_char_p]
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("prfdd %s failed: %s (code %d)" % (what, lib().prfdd_error_string(rc).decode(), rc))


def ladder(N, r):
    out = [N]
    while out[-1] > 1:
        out.append(max(out[-1] - r, 1))
    return out


def mesh_generate_box(directory, dim, nel, N, num_procs=1, eps=0.0, reduction=None, only_rank=-1):
    """Writes the mesh for degree N (and, if reduction is given, for every degree of the ladder); only_rank >= 0: that rank's files only."""
    if isinstance(nel, int):
        nel = (nel,) * dim
    nel3 = (C.c_int * 3)(*(list(nel) + [1] * (3 - len(nel))))
    degrees = ladder(N, reduction) if reduction else [N]
    for n in degrees:
        check(lib().prfdd_mesh_generate_box_rank(directory.encode(), C.c_int(dim), nel3, C.c_int(n), C.c_int(num_procs), C.c_double(eps), C.c_int(only_rank)), "mesh_generate_box")


class Solver:
    """poisson.cpp's run_simulation() behind a handle: create -> setup_problem -> solve."""

    def __init__(self, directory, stream=0, **kw):
        L = lib()
        self.opt = Options()
        L.prfdd_options_default(C.byref(self.opt))
        self._uid = None
        for k, v in kw.items():
            if k == "nccl_unique_id":
                if v is not None:
                    self._uid = C.create_string_buffer(bytes(v), 128)
                    self.opt.nccl_unique_id = C.cast(self._uid, C.c_void_p)
            else:
                if not hasattr(self.opt, k):
                    raise TypeError("unknown option " + k)
                setattr(self.opt, k, v)
        self.h = C.c_void_p()
        check(L.prfdd_solver_create(C.byref(self.h), directory.encode(), C.byref(self.opt), C.c_void_p(stream)), "solver_create")

    def close(self):
        if self.h:
            lib().prfdd_solver_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def query(self, name):
        return int(lib().prfdd_solver_query(self.h, Q[name]))

    def setup_problem(self, function_id=4):
        check(lib().prfdd_solver_setup_problem(self.h, C.c_int(function_id)), "setup_problem")

    def _hist(self):
        return np.zeros(1024), C.c_int(0), C.c_int(0)

    def solve(self, solver_id=0):
        hist, nit, hl = self._hist()
        check(lib().prfdd_solver_solve(self.h, C.c_int(solver_id), C.byref(nit), hist.ctypes.data_as(C.c_void_p), C.c_int(hist.size), C.byref(hl)), "solve")
        return nit.value, hist[:hl.value].copy()

    def solve_host(self, f_host, u_host, solver_id=0):
        hist, nit, hl = self._hist()
        check(lib().prfdd_solver_solve_host(self.h, C.c_int(solver_id), f_host.ctypes.data_as(C.c_void_p), u_host.ctypes.data_as(C.c_void_p),
                                            C.byref(nit), hist.ctypes.data_as(C.c_void_p), C.c_int(hist.size), C.byref(hl)), "solve_host")
        return nit.value, hist[:hl.value].copy()

    def get_array(self, name, count=None):
        dt = np.dtype(A_DTYPE.get(name, np.float64))
        if count is None:
            probe = int(lib().prfdd_solver_get_array(self.h, A[name], None, C.c_longlong(0)))
            if probe == 0:
                return np.zeros(0, dtype=dt)
            if probe > 0 or probe == -1:
                raise RuntimeError("get_array(%s) failed" % name)
            count = (-probe) // dt.itemsize
        out = np.zeros(count, dtype=dt)
        n = int(lib().prfdd_solver_get_array(self.h, A[name], out.ctypes.data_as(C.c_void_p), C.c_longlong(out.nbytes)))
        if n < 0:
            raise RuntimeError("get_array(%s) failed (%d)" % (name, n))
        return out[:n]

    def apply(self, what, x, out_len=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(out_len if out_len is not None else x.size)
        check(lib().prfdd_solver_apply(self.h, C.c_int(APPLY[what]), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)), "apply " + what)
        return out

    def output(self, name):
        """u_star, f, u of this rank's elements -> <name>_<rank>.vtk (poisson.cpp:233-235)"""
        check(lib().prfdd_solver_output(self.h, name.encode()), "output")

    def profile_vcycle(self, reps=5):
        """per-launch table of this rank's AMG V-cycle (prfdd_solver_profile_vcycle)"""
        buf = C.create_string_buffer(1 << 16)
        check(lib().prfdd_solver_profile_vcycle(self.h, C.c_int(reps), buf, C.c_int(len(buf))), "profile_vcycle")
        return buf.value.decode()

    def timer(self, key):
        return float(lib().prfdd_solver_timer_total(self.h, key.encode()))
