#!/usr/bin/env python
"""
bench.py -- headline benchmark of the PR-FDD preconditioned CG hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPU (oracle port)

metric   FP64 PCG GDOF*iter/s  = (global unique GLL nodes x outer PCG iterations) / solve time / 1e9
step     one complete preconditioned CG solve (flexible CG, relative residual 1e-8) of the synthetic problem
         f = A u*, u* = glibc rand() stream (the reference's function_id 4), through the C ABI
workload N = 1: BASELINE configs[1] -- 3D SEM Poisson, 16^3 hex box mesh, polynomial degree 7, FP64
         N > 1, --scaling weak (default): 16^3 elements per GPU, block partition (N = 8 is configs[2]: 32^3 on 2x2x2)
         --scaling strong: configs[2] as written -- the FIXED 32^3 mesh on N = 1, 2, 4, 8 GPUs
value    device-resident f (only the residual norm is read by the host each iteration)
e2e      the same solve through prfdd_solver_solve_host: f copied host->device and u device->host inside
         the timed region, every step
parity   the solve's iteration count and residual history against the ORACLE's solve of the same mesh and partition,
         computed offline (tests/golden/bench_histories.json; generator tests/golden/make_bench_histories.py), at every N;
         a mismatch makes the run fail (exit code 3) after the JSON line is printed
One JSON line on stdout (rank 0).

--impl reference: the reference cannot be built (OCCA / HYPRE / GSLib / MPI / gfortran absent); its algorithm is timed as the
oracle port -- the OKL kernels as the plain loops OCCA Serial runs, with the @outer loops on all host threads (the OCCA OpenMP
analogue) -- on the SAME workload as the B200 arm at N = 1 (configs[1] in full).  A step of that arm is ONE outer PCG iteration
(the metric is per iteration), so that K steps stay within minutes; a single-thread leg of one solve is reported beside it.
With --gpus N > 1 the line names the B200 arm's workload at that N and times a bounded sample of it: one 16^3 block (1/N of the
weak-scaled mesh's elements) as a one-process problem (`config.sample_of_workload`); a full-size oracle solve needs up to an hour of set-up.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fp64_pcg_gdof_iter_per_s"
UNIT = "GDOF*iter/s"
TOL = 1.0e-8            # BASELINE.json: "iters to 1e-8" (the reference's own default is 1e-7, domain.hpp:118)
N_DEG, REDUCTION = 7, 3  # degree ladder 7, 4, 1 (profile.sh:5-11)
NEL_PER_GPU = 16
NEL_STRONG = 32
HISTORY_RTOL = 1.0e-9   # residual history against the oracle's (the bar of tests/test_gpu_subdomain.py)
GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_histories.json")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def layout(nranks):
    p = [1, 1, 1]
    d, r = 0, nranks
    while r > 1:
        p[d % 3] *= 2
        r //= 2
        d += 1
    return p


def mesh_of(world, scaling):
    P3 = layout(world)
    if scaling == "strong":
        return P3, [NEL_STRONG] * 3
    return P3, [NEL_PER_GPU * P3[0], NEL_PER_GPU * P3[1], NEL_PER_GPU * P3[2]]


def golden_record(nel, ranks, eps, coarsening):
    """the oracle's solve of this mesh / partition (None if it was never generated)"""
    try:
        recs = json.load(open(GOLDEN))
    except Exception:
        return None
    for r in recs:
        if (list(r["nel"]) == list(nel) and r["N"] == N_DEG and r["r"] == REDUCTION and abs(r["eps"] - eps) < 1e-15 and r["ranks"] == ranks
                and abs(r["tolerance"] - TOL) < 1e-20 and r.get("coarsening", "hmis") == coarsening):
            return r
    return None


def parity_block(hist, nel, ranks, eps, coarsening, rel_error, precision="double"):
    rec = golden_record(nel, ranks, eps, coarsening)
    rtol = HISTORY_RTOL if precision == "double" else 1.0e-4      # FP32 V-cycle: the history follows the FP64 oracle's to FP32 accuracy
    out = {"checked": False, "iterations": len(hist) - 1, "rel_error_vs_exact": rel_error, "history": [float(h) for h in hist]}
    # the manufactured solution is known: the converged iterate must be within the solve tolerance's reach of it at every size
    out["exact_solution_ok"] = bool(rel_error is not None and rel_error < 1.0e-6)
    if rec is None:
        out["why_unchecked"] = "no oracle record for this mesh / partition / coarsening in tests/golden/bench_histories.json"
        out["ok"] = out["exact_solution_ok"]
        return out
    gh = np.array(rec["history"])
    out["checked"] = True
    out["golden"] = "tests/golden/bench_histories.json:%s (oracle, %s coarsening, %d simulated rank(s); oracle solve %.0f s + set-up %.0f s on the build host)" % (
        rec["name"], rec.get("coarsening", "hmis"), rec["ranks"], rec.get("oracle_solve_s", 0), rec.get("oracle_setup_s", 0))
    out["iterations_oracle"] = len(gh) - 1
    same_len = len(gh) == len(hist)
    out["history_max_rel_diff"] = float(np.max(np.abs(np.array(hist) - gh) / gh)) if same_len else None
    out["history_rtol"] = rtol
    out["ok"] = bool(same_len and out["history_max_rel_diff"] <= rtol and out["exact_solution_ok"])
    return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU legs (oracle port)
# ---------------------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def set_omp_threads(n):
    """the OpenMP build of the oracle kernels is loaded once; the thread count is switched at run time"""
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(int(n)))
        return True
    except Exception:
        return False


def oracle_problem(nel, threads):
    from oracle import capi
    capi.set_threads(threads)
    from oracle import meshgen, domain as od, subdomain as osub
    d = tempfile.mkdtemp(prefix="prfdd_cpu_")
    t0 = time.perf_counter()
    for n in osub.ladder(N_DEG, REDUCTION):
        meshgen.generate(d, 3, nel, n, nranks=1, eps=0.0)
    W = od.DomainWorld(d, N_DEG, 1)
    W.tolerance = TOL
    Sd = osub.SubdomainWorld(W, d, N_DEG, REDUCTION)
    us = W.initial_function(4)
    f = W.new_vector(); W.stiffness_matrix(f, us)
    nodes = int(np.prod([n * N_DEG + 1 for n in (nel if isinstance(nel, (list, tuple)) else [nel] * 3)]))
    return W, Sd, f, nodes, time.perf_counter() - t0


def osub_ladder(N, r):
    out = [N]
    while out[-1] > 1:
        out.append(max(out[-1] - r, 1))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    os.environ.setdefault("OMP_PROC_BIND", "false")
    cores = host_threads()
    nel = [NEL_PER_GPU] * 3                                   # configs[1], the B200 arm's N = 1 workload, in full
    W, Sd, f, nodes, setup_s = oracle_problem(nel, max(cores, 2))

    def iterations(count):
        """`count` outer PCG iterations: whole solves, the last one cut by max_iterations"""
        done = 0
        while done < count:
            u = W.new_vector()
            W.flexible_conjugate_gradient(u, f, Sd, max_iterations=count - done)
            done += len(W.history) - 1
        return done

    # calibration (also the warm-up): one full solve with every host thread and one with a single thread; the timed steps run in the
    # faster of the two configurations, both are reported
    legs = {}
    full_iters = None
    for threads in ([cores, 1] if cores > 1 else [1]):
        if not set_omp_threads(threads):
            continue
        t0 = time.perf_counter()
        u = W.new_vector(); W.flexible_conjugate_gradient(u, f, Sd)
        dt = time.perf_counter() - t0
        full_iters = len(W.history) - 1
        legs[threads] = {"value": nodes * full_iters / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "one full solve (%d iterations) of the same workload on %d host thread(s) (%s analogue)" % (full_iters, threads, "OCCA OpenMP" if threads > 1 else "OCCA Serial")}
    best = max(legs, key=lambda k: legs[k]["value"]) if legs else 1
    set_omp_threads(best)
    if args.warmup > full_iters:
        iterations(args.warmup - full_iters)
    t0 = time.perf_counter()
    iters = iterations(args.steps)
    dt = time.perf_counter() - t0
    value = nodes * iters / dt / 1e9
    sample = ("3D 16^3 hex box, N=7 (%d nodes), ladder 7/4/1 -- configs[1] in full; a step = ONE outer PCG iteration of the PR-FDD solve to 1e-8 (%d iterations per solve); "
              "%d iterations timed on %d of %d host threads (the faster of the all-threads and single-thread configurations; the oracle's @outer loops run on the threads: "
              "OCCA OpenMP / Serial analogue); set-up (%.0f s) excluded" % (nodes, full_iters, iters, best, cores, setup_s))
    # the B200 arm's workload at this GPU count (same string as its `config.workload`); the CPU arm times a bounded sample of it: at N = 1
    # the workload itself, at N > 1 one 16^3 block of it (1/N of the elements of the weak-scaled mesh) solved as a one-rank problem
    P3, arm_nel = mesh_of(max(args.gpus, 1), args.scaling)
    workload = ("3D SEM Poisson, %dx%dx%d hex box mesh, N=%d, FP64, PR-FDD preconditioned flexible CG (ladder %s, overlaps 1/1, inner GMRES(4), 1 V-cycle, Chebyshev order 2)"
                % (tuple(arm_nel) + (N_DEG, "/".join(str(x) for x in osub_ladder(N_DEG, REDUCTION)))))
    whole = list(arm_nel) == list(nel)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(iters, 1), "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload,
                       "tolerance": TOL, "global_nodes": nodes, "iterations_per_solve": full_iters, "step": "one outer PCG iteration (bounded sample of the solve)",
                       "sample_of_workload": "the workload in full (16^3 elements, one process)" if whole else
                                             "one 16x16x16 block of the %dx%dx%d mesh (1/%d of its elements) solved as a one-process problem with all host threads: a full-size oracle solve takes "
                                             "up to an hour of set-up (tests/golden/bench_histories.json), and the N-rank composite problems are not simulated on the CPU"
                                             % (tuple(arm_nel) + (int(np.prod(arm_nel)) // int(np.prod(nel)),)),
                       "partition": "1 process",
                       "setup_s": setup_s, "host_threads_available": cores},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": best, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "cpu_legs": [legs[k] for k in sorted(legs, reverse=True)]}
    print(json.dumps(line), flush=True)


def cpu_baseline(nel):
    """bounded sample beside the B200 arm: the oracle port on a mesh of nel^3 elements, all host threads"""
    os.environ.setdefault("OMP_WAIT_POLICY", "passive")
    os.environ.setdefault("OMP_PROC_BIND", "false")
    cores = host_threads()
    W, Sd, f, nodes, setup_s = oracle_problem(nel, max(cores, 2))
    # one warm solve per thread configuration; the faster one is timed
    rate = {}
    for threads in ([cores, 1] if cores > 1 else [1]):
        if set_omp_threads(threads):
            t0 = time.perf_counter()
            u = W.new_vector(); W.flexible_conjugate_gradient(u, f, Sd)
            rate[threads] = (len(W.history) - 1) / (time.perf_counter() - t0)
    cores = max(rate, key=rate.get) if rate else 1
    set_omp_threads(cores)
    t0 = time.perf_counter(); iters = 0; reps = 0
    while time.perf_counter() - t0 < 10.0 and reps < 5:
        u = W.new_vector(); W.flexible_conjugate_gradient(u, f, Sd)
        iters += len(W.history) - 1; reps += 1
    dt = time.perf_counter() - t0
    return {"value": nodes * iters / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle port (C kernels = the OKL loops, @outer loops on %d host threads) on the box's host CPU: 3D %d^3 hex box (1/%d of the workload's elements), N=7 (%d nodes), "
                      "%d full PR-FDD PCG solves to 1e-8, set-up (%.0f s) excluded; the full workload is what `--impl reference` runs" % (cores, nel, (NEL_PER_GPU // nel) ** 3, nodes, reps, setup_s)}


# ---------------------------------------------------------------------------------------------------------------------
# kernel rooflines
# ---------------------------------------------------------------------------------------------------------------------
def ncu_traffic(o):
    """DRAM bytes per launch (read + write) of the roofline kernel from the committed `ncu --set full` capture, averaged over the
    two levels bench alternates -- only if the capture is of these very matrices (same rows / nnz), else null"""
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", name)))["cheby_step"]
            l0, l1 = t["level0"], t["level1"]
            if (l0["rows"], l0["nnz"], l1["rows"], l1["nnz"]) == (int(o[2]), int(o[3]), int(o[4]), int(o[5])):
                return 0.5 * (l0["dram_bytes"] + l1["dram_bytes"])
        except Exception:
            pass
    return None


def operator_traffic(E, n):
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))["k_ax3d_bulk"]
        return float(t["dram_bytes"]) if (t["elements"], t["n"]) == (E, n) else None
    except Exception:
        return None


def operator_roofline(pr, torch, stream, peaks, peaks_kind):
    """average launch duration of the SEM operator kernel (the dominant kernel outside the V-cycle), CUDA events on the launching
    stream, three rotating operand sets (3 x 134 MB > L2)"""
    L = pr.lib()
    E, n = NEL_PER_GPU ** 3, N_DEG + 1
    P = E * n ** 3
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    z = np.zeros(n); w = np.zeros(n); D = np.zeros(n * n)
    L.prfdd_zwgll(z.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), C.c_int(n))
    L.prfdd_dgll(D.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), C.c_int(n))
    Dd = torch.from_numpy(D).cuda()
    sets = []
    for _ in range(3):
        uu = torch.rand(P, dtype=torch.float64, device="cuda", generator=g)
        GG = [torch.rand(P, dtype=torch.float64, device="cuda", generator=g) for _ in range(6)]
        sets.append((uu, GG, (C.c_void_p * 6)(*[t.data_ptr() for t in GG]), torch.empty(P, dtype=torch.float64, device="cuda")))
    sh = C.c_void_p(stream.cuda_stream)
    times = []
    reps = 12
    with torch.cuda.stream(stream):
        for it in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in range(reps):
                uu, GG, gp, out = sets[r % 3]
                rc = L.prfdd_stiffness_matrix_hd(C.c_void_p(out.data_ptr()), C.c_void_p(uu.data_ptr()), C.c_void_p(Dd.data_ptr()), D.ctypes.data_as(C.c_void_p), gp,
                                                 C.c_int(E), C.c_int(n), C.c_int(3), sh)
                assert rc == 0
            e1.record(stream)
            stream.synchronize()
            if it >= 2:
                times.append(e0.elapsed_time(e1) / reps)
    t_ms = float(np.mean(times))
    bytes_alg = 64.0 * P                                  # u 8 + six G 48 + Au 8 per point
    achieved = bytes_alg / (t_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "%s (prfdd_stiffness_matrix_hd, %d elements, n=%d)" % ("k_ax3d_bulk<8>" if n == 8 else "k_ax3d_bulk<6>" if n == 6 else "k_ax3d_big<%d>" % n if n > 10 else "k_ax3d<%d>" % n, E, n), "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "traffic": operator_traffic(E, n), "peak_source": peaks_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peaks_kind == "measured" else "fallback 6.65 TB/s",
            "avg_launch_ms": t_ms, "algorithmic_bytes_per_launch": bytes_alg}


PHASE_KEYS = ["domain.operator_application", "domain.inner_products", "domain.residual_norm", "domain.vector_operations", "subdomain.stitching",
              "subdomain.tree_construction.gpu_to_gpu", "subdomain.tree_construction.subdomain", "subdomain.tree_construction.assemble_coarse",
              "subdomain.tree_construction.superdomain", "subdomain.tree_exchange.subdomain", "subdomain.tree_exchange.superdomain",
              "subdomain.preconditioner", "subdomain.preconditioner.assemble_subdomain", "subdomain.preconditioner.assemble_composite",
              "subdomain.preconditioner.down_leg_gpu", "subdomain.preconditioner.unassemble_composite", "subdomain.preconditioner.unassemble_subdomain",
              "subdomain.operator_application", "subdomain.inner_products", "subdomain.residual_norm", "subdomain.vector_operations"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="weak: 16^3 elements per GPU; strong: the fixed 32^3 mesh (configs[2] as written)")
    ap.add_argument("--cpu-nel", type=int, default=8, help="elements per side of the bounded CPU sample beside the B200 arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eps", type=float, default=0.0, help="mesh deformation (the reference's profiling meshes are Kershaw eps = 0.3, profile.sh:5-11)")
    ap.add_argument("--coarsening", default="hmis", choices=["hmis", "pmis"])
    ap.add_argument("--amg-precision", default="double", choices=["double", "float"], help="float: the FP32 V-cycle (`Float float`, AMG/config.hpp:4; SURVEY 8f N3); the headline runs double")
    ap.add_argument("--degree", type=int, default=7, help="polynomial degree N (headline: 7); other values are extra records (configs[3], configs[4]), not the headline")
    ap.add_argument("--reduction", type=int, default=3, help="degree reduction per ladder step (headline: 3 -> ladder 7, 4, 1)")
    ap.add_argument("--nel-per-gpu", type=int, default=16, help="elements per side of a rank's block in weak scaling (headline: 16)")
    ap.add_argument("--device-outer-loop", action="store_true", help="one rank: run the outer solve as one CUDA graph with a device-side loop test (prfdd_options.device_outer_loop)")
    ap.add_argument("--phases", action="store_true", help="add the fenced per-phase table (reference Timer keys) of one extra solve")
    args = ap.parse_args()
    global N_DEG, REDUCTION, NEL_PER_GPU
    N_DEG, REDUCTION, NEL_PER_GPU = args.degree, args.reduction, args.nel_per_gpu
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import polynomial_reduction_with_full_domain_decomposition_preconditioner_b200 as pr

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    L = pr.lib()
    L.prfdd_algorithmic_bytes.restype = C.c_double
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # one ncclUniqueId for the library's own communicator, broadcast with the torch process group
        holder = [None]
        if rank == 0:
            nccl = C.CDLL("libnccl.so.2")
            buf = C.create_string_buffer(128)
            assert nccl.ncclGetUniqueId(buf) == 0
            holder[0] = bytes(buf.raw)
        dist.broadcast_object_list(holder, src=0)
        uid = holder[0]

    # synthetic mesh in the reference's on-disk format: every rank's files, written once by rank 0
    P3, nel = mesh_of(world, args.scaling)
    if world > 1:
        holder = [tempfile.mkdtemp(prefix="prfdd_bench_") if rank == 0 else None]
        dist.broadcast_object_list(holder, src=0)
        mesh_dir = holder[0]
    else:
        mesh_dir = tempfile.mkdtemp(prefix="prfdd_bench_")
    use_pc = 0 if os.environ.get("PRFDD_BENCH_NO_PC") else 1
    t_setup0 = time.perf_counter()
    # every rank writes its own part of the mesh
    pr.mesh_generate_box(mesh_dir, 3, tuple(nel), N_DEG, world, args.eps, reduction=REDUCTION if use_pc else None, only_rank=rank if world > 1 else -1)
    if world > 1:
        dist.barrier()
    mesh_s = time.perf_counter() - t_setup0

    stream = torch.cuda.Stream()
    S = pr.Solver(mesh_dir, stream=stream.cuda_stream, poly_degree=N_DEG, poly_reduction=REDUCTION, use_preconditioner=use_pc,
                  outer_tolerance=TOL, proc_id=rank, num_procs=world, nccl_unique_id=uid, amg_coarsening=0 if args.coarsening == "pmis" else 1, amg_precision=1 if args.amg_precision == "float" else 0, device_outer_loop=1 if args.device_outer_loop else 0)
    S.setup_problem(4)
    stream.synchronize()
    setup_s = time.perf_counter() - t_setup0
    nodes = S.query("NUM_GLOBAL_NODES")
    P = S.query("NUM_LOCAL_POINTS")
    f_host = torch.from_numpy(S.get_array("F")).pin_memory()
    u_host = torch.empty(P, dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        L.prfdd_launch_count_reset()
        L.prfdd_algorithmic_bytes_reset()
        e0.record(stream)
        iters = 0
        for _ in range(steps):
            iters += fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = int(L.prfdd_launch_count())
        alg_bytes = float(L.prfdd_algorithmic_bytes())
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            b = torch.tensor([alg_bytes], dtype=torch.float64, device="cuda")
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            alg_bytes = float(b.item())
        return ms, iters, launches, alg_bytes

    def step_dev():
        nit, hist = S.solve(0)
        return len(hist) - 1

    def step_host():
        nit, hist = S.solve_host(f_host.numpy(), u_host.numpy(), 0)
        return len(hist) - 1

    for _ in range(args.warmup):
        step_dev()
    for _ in range(2):
        step_host()
    sampler = ClockSampler(local)
    sampler.start()
    # cudaProfilerStart/Stop around the timed steps: `ncu --profile-from-start off ... python bench.py ...` then lists exactly the
    # launches of the timed region (set-up alone makes ~60k launches); a no-op without a profiler
    torch.cuda.cudart().cudaProfilerStart()
    ms, iters, launches, alg_bytes = timed(step_dev, args.steps)
    ms_e2e, iters_e2e, _, _ = timed(step_host, args.steps)
    torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop()
    value = nodes * iters / (ms * 1e-3) / 1e9
    e2e = nodes * iters_e2e / (ms_e2e * 1e-3) / 1e9
    nit, hist = S.solve(0)
    # relative error against the manufactured solution over ALL ranks' points
    us, ug = S.get_array("U_STAR"), S.get_array("U")
    num, den = float(np.sum((ug - us) ** 2)), float(np.sum(us ** 2))
    if world > 1:
        t = torch.tensor([num, den], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        num, den = float(t[0].item()), float(t[1].item())
    err = float(np.sqrt(num / den))

    phases = None
    if args.phases:
        # the reference's Timer keys; enabling the Timer fences every phase and disables graph replay: read the shares
        S.timer("__enable__")
        S.solve(0)
        vals = [max(S.timer(k), 0.0) for k in PHASE_KEYS]
        S.timer("__disable__")
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        phases = {k: round(1e3 * float(v), 3) for k, v in zip(PHASE_KEYS, t.tolist()) if v > 0}
        phases["_note"] = "ms of one fenced solve, max over ranks (fencing and the absence of graph replay inflate small phases)"

    vprof = None
    if args.phases and use_pc:
        mine = S.profile_vcycle(5)
        if world > 1:
            allp = [None] * world
            dist.all_gather_object(allp, mine)
            vprof = {"rank0": allp[0].splitlines(), "rank%d" % (world - 1): allp[-1].splitlines()}
        else:
            vprof = {"rank0": mine.splitlines()}
    amg_rows = [int(x) for x in S.get_array("AMG_LEVEL_ROWS")] if use_pc else []
    amg_nnz = [int(x) for x in S.get_array("AMG_LEVEL_NNZ")] if use_pc else []
    sizes = None
    if use_pc:
        mine = [S.query("SUB_NUM_POINTS"), S.query("NUM_DOFS"), S.query("AMG_NUM_LEVELS"), amg_rows[0], int(sum(amg_nnz))]
        if world > 1:
            allv = [None] * world
            dist.all_gather_object(allv, mine)
        else:
            allv = [mine]
        sizes = {"region_points_per_rank": [a[0] for a in allv], "composite_dofs_per_rank": [a[1] for a in allv], "amg_levels_per_rank": [a[2] for a in allv],
                 "amg_total_nnz_per_rank": [a[4] for a in allv]}

    rc = 0
    if rank == 0:
        peaks, kind = measured_peaks()
        roof_op = operator_roofline(pr, torch, stream, peaks, kind)
        roof = roof_op
        if use_pc:
            o = (C.c_double * 6)()
            if L.prfdd_solver_time_spmv(S.h, C.c_int(40), o) == 0:
                ach = o[1] / (o[0] * 1e-3) / 1e9
                roof = {"bound": "hbm", "kernel": "sliced SpMV (k_spmv_sell / k_spmv_sell_window) + Chebyshev epilogue (prfdd_csrm_cheby_step) on AMG levels 0 and 1 of the low-order FEM hierarchy, alternating "
                        "(level 0: %d rows, %d nnz; level 1: %d rows, %d nnz) -- the SpMV family is the dominant kernel of the solve (profiles/)" % (o[2], o[3], o[4], o[5]),
                        "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": ncu_traffic(o),
                        "peak_source": roof_op["peak_source"], "avg_launch_ms": o[0], "algorithmic_bytes_per_launch": o[1],
                        "algorithmic_bytes": "12 B per entry (col 4 + val 8) + 4 B per row pointer + 32 B per row (t_in, ds, r read, t_out written); FP32 V-cycle: 8 B per entry, 16 B per row"}
        # whole solve: sum of the algorithmic bytes of every kernel launched in the timed region (each launcher reports its own:
        # prfdd_algorithmic_bytes) over the device time, per GPU
        per_gpu_gbs = alg_bytes / world / (ms * 1e-3) / 1e9
        whole = {"bound": "hbm", "algorithmic_bytes_per_solve": alg_bytes / args.steps, "achieved": per_gpu_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s per GPU",
                 "frac": per_gpu_gbs / peaks["hbm_gbs"], "note": "sum over all kernels of a solve of each kernel's algorithmic bytes (SURVEY 8d) / solve time / peak"}
        parity = parity_block(hist, nel, world, args.eps, args.coarsening, err, args.amg_precision) if use_pc else {"checked": False, "ok": True, "why_unchecked": "no preconditioner"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args.cpu_nel)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "3D SEM Poisson, %dx%dx%d hex box mesh, N=%d, FP64, PR-FDD preconditioned flexible CG (ladder %s, overlaps 1/1, inner GMRES(4), 1 V-cycle, Chebyshev order 2)"
                           % (tuple(nel) + (N_DEG, "/".join(str(x) for x in pr.ladder(N_DEG, REDUCTION))))
                           if use_pc else "3D SEM Poisson, %dx%dx%d hex box mesh, N=%d, FP64, flexible CG WITHOUT the PR-FDD preconditioner (PRFDD_BENCH_NO_PC set)" % (tuple(nel) + (N_DEG,)),
                           "tolerance": TOL, "global_nodes": nodes, "iterations_per_solve": iters // args.steps, "l2": "working set (geometry 96 MiB + vectors + AMG hierarchy) exceeds the 126 MB L2",
                           "partition": "%dx%dx%d blocks of %dx%dx%d elements" % (tuple(P3) + tuple(n // p for n, p in zip(nel, P3))), "mesh_deformation_eps": args.eps,
                           "amg_coarsening": args.coarsening, "amg_precision": args.amg_precision, "time_to_solution_ms": ms / args.steps, "setup_s": setup_s, "mesh_generation_s": mesh_s, "rel_error_vs_exact": err,
                           "amg_level_rows_rank0": amg_rows, "amg_level_nnz_rank0": amg_nnz},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * P * world, "d2h_bytes_per_step": 8 * P * world, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roof, "roofline_sem_operator": roof_op, "roofline_whole_solve": whole, "parity": parity}
        if sizes:
            line["sizes"] = sizes
        if phases:
            line["phases"] = phases
        if vprof:
            line["vcycle_profile"] = {"columns": "level what rows nnz us algorithmic_MB GB/s (every launch of one V-cycle between CUDA events, average of 5 cycles)", **vprof}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
        if not parity.get("ok", False):
            print("bench.py: PARITY FAILED against the oracle record: %s" % json.dumps({k: v for k, v in parity.items() if k != "history"}), file=sys.stderr, flush=True)
            rc = 3
    S.close()
    if world > 1:
        dist.destroy_process_group()
    if rc:
        sys.exit(rc)


if __name__ == "__main__":
    main()
