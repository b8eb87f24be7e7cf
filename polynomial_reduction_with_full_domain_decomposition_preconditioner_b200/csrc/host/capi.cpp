// capi.cpp -- solver-level C ABI: the reference's run_simulation() sequence (poisson.cpp:150-251)
// behind an opaque handle.  See include/prfdd_b200.h.
#include "config.hpp"
#include "domain.hpp"
#include "subdomain.hpp"
#include "../../../include/prfdd_b200.h"
#include <cstring>
#include <map>
#include <memory>
#include <string>

using namespace prfdd_host;

struct prfdd_solver
{
    prfdd_options opt;
    std::string directory;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int device_id = 0;
    int dim_ = 0;
    std::map<int, std::unique_ptr<Domain<STYPE>>> domains; // ladder of Domain objects (poisson.cpp:172-199)
    std::unique_ptr<Subdomain<PTYPE>> subdomain;
    Domain<STYPE> *domain = nullptr;
    dev::memory u_star, f, u;
    dev::memory apply_in, apply_out; // persistent operands of prfdd_solver_apply: a stable key for the preconditioner's graph cache
    double *pin_in = nullptr, *pin_out = nullptr;
    Comm comm;
    Timer<double> tmr;
    int function_id = 4;

    void activate()
    {
        // the host classes see process-wide globals, like the reference (config.hpp:48-65)
        cudaSetDevice(device_id);
        prfdd_host::dim = dim_;
        prfdd_host::proc_id = opt.proc_id;
        prfdd_host::num_procs = opt.num_procs;
        prfdd_host::verbose = opt.verbose;
        prfdd_host::device.setup(device_id, stream);
        prfdd_host::comm_world = comm;
        prfdd_host::timer = tmr;
    }
    void deactivate()
    {
        dim_ = prfdd_host::dim;
        tmr = prfdd_host::timer;
    }
};

namespace
{
template <class F>
int guarded(prfdd_solver *s, F fn)
{
    try
    {
        if (s) s->activate();
        int rc = fn();
        if (s) s->deactivate();
        return rc;
    }
    catch (const std::exception &ex)
    {
        fprintf(stderr, "prfdd: %s\n", ex.what());
        return -1;
    }
}
} // namespace

extern "C" {

void prfdd_options_default(prfdd_options *o)
{
    memset(o, 0, sizeof(*o));
    o->poly_degree = 7;
    o->poly_reduction = 3;
    o->subdomain_overlap = 1;
    o->superdomain_overlap = 1;
    o->use_preconditioner = 1;
    o->preconditioner_type = 1;
    o->inner_num_vectors = 4;
    o->inner_max_iterations = 4;
    o->num_vcycles = 1;
    o->cheby_order = 2;
    o->use_cuda_graph = 1;
    o->proc_id = 0;
    o->num_procs = 1;
    o->nccl_unique_id = nullptr;
    o->outer_tolerance = 1.0e-7;
    o->inner_tolerance = 1.0e-12;
    o->outer_max_iterations = 500;
    o->outer_num_vectors = 20;
    o->verbose = 0;
    o->amg_coarsening = -1;
    o->amg_precision = 0;
    o->device_outer_loop = 0;
}

int prfdd_solver_create(prfdd_solver **out, const char *directory, const prfdd_options *opt, prfdd_stream_t stream)
{
    prfdd_solver *s = new prfdd_solver();
    s->opt = *opt;
    s->directory = directory;
    s->stream = (cudaStream_t)stream;
    cudaGetDevice(&s->device_id);
    if (s->stream == nullptr)
    {
        // the legacy default stream cannot be captured into a CUDA graph: run on an own stream instead
        if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) { delete s; return -1; }
        s->own_stream = true;
    }
    int rc = guarded(nullptr, [&]() {
        s->comm.init(opt->proc_id, opt->num_procs, opt->nccl_unique_id, s->stream);
        s->tmr.stream = s->stream;
        s->tmr.initialize();
        s->activate();

        // ladder of domains (poisson.cpp:172-199)
        rstdout("Running simulation with:\n");
        rstdout("- Directory: \"%s\"\n", directory);
        rstdout("- Polynomial degree: \"%d\"\n", opt->poly_degree);
        rstdout("- Polynomial reduction: \"%d\"\n", opt->poly_reduction);
        rstdout("- Subdomain overlap: \"%d\"\n", opt->subdomain_overlap);
        rstdout("- Superdomain overlap: \"%d\"\n", opt->superdomain_overlap);
        int N = opt->poly_degree;
        rstdout("\nSetting up domain \"N = %d\" object...\n", N);
        setup_mark(nullptr);
        s->domains[N].reset(new Domain<STYPE>());
        s->domains[N]->num_vectors = opt->outer_num_vectors;
        s->domains[N]->initialize(directory, N, true);
        s->domain = s->domains[N].get();
        s->domain->max_iterations = opt->outer_max_iterations;
        s->domain->tolerance = opt->outer_tolerance;
        s->domain->preconditioner_type = opt->preconditioner_type;
        s->domain->use_preconditioner = opt->use_preconditioner != 0;
        {
            static const char *env = getenv("PRFDD_DEVICE_OUTER_LOOP"); // experiment knob: overrides the option
            const int want = env ? atoi(env) : opt->device_outer_loop;
            // one rank only: with NCCL collectives inside the WHILE body the 2-rank run hung (measured, killed by its timeout)
            s->domain->device_outer_loop = want != 0 && opt->use_cuda_graph != 0 && opt->num_procs == 1;
        }
        if (opt->use_preconditioner)
        {
            int level = N;
            while (level > 1)
            {
                level -= opt->poly_reduction;
                if (level < 1) level = 1;
                rstdout("Setting up domain \"N = %d\" object...\n", level);
                s->domains[level].reset(new Domain<STYPE>());
                s->domains[level]->initialize(directory, level, false);
            }
            rstdout("Setting up subdomain object...\n");
            setup_mark("Domain objects (reader, Q, halo lists)");
            s->subdomain.reset(new Subdomain<PTYPE>(s->domains, N, opt->poly_reduction, opt->subdomain_overlap, opt->superdomain_overlap, *opt));
        }
        else
        {
            s->subdomain.reset(new Subdomain<PTYPE>()); // never applied
        }
        int P = s->domain->num_local_points;
        s->u_star = device.malloc<STYPE>(P);
        s->f = device.malloc<STYPE>(P);
        s->u = device.malloc<STYPE>(P);
        dev::check(cudaMallocHost((void **)&s->pin_in, sizeof(double) * P), "pinned in");
        dev::check(cudaMallocHost((void **)&s->pin_out, sizeof(double) * P), "pinned out");
        device.finish();
        s->deactivate();
        return 0;
    });
    if (rc != 0)
    {
        delete s;
        return rc;
    }
    *out = s;
    return 0;
}

int prfdd_solver_destroy(prfdd_solver *s)
{
    if (!s) return 0;
    guarded(s, [&]() {
        device.finish();
        s->subdomain.reset();
        s->domains.clear();
        s->u_star.free(); s->f.free(); s->u.free();
        if (s->pin_in) cudaFreeHost(s->pin_in);
        if (s->pin_out) cudaFreeHost(s->pin_out);
        s->comm.finalize();
        return 0;
    });
    if (s->own_stream) cudaStreamDestroy(s->stream);
    delete s;
    return 0;
}

int prfdd_solver_setup_problem(prfdd_solver *s, int function_id)
{
    return guarded(s, [&]() {
        s->function_id = function_id;
        rstdout("\nSetting up exact function...\n");
        s->domain->initial_function(s->u_star, function_id);  // poisson.cpp:211-213
        rstdout("Setting up right-hand-side...\n");
        s->domain->stiffness_matrix(s->f, s->u_star);           // poisson.cpp:218-219
        device.finish();
        return 0;
    });
}

static int do_solve(prfdd_solver *s, int solver_id, int *num_iterations, double *history, int history_cap, int *history_len)
{
    rstdout("Solving Poisson problem...\n");
    if (solver_id == 0)
        s->domain->flexible_conjugate_gradient(s->u, s->f, *s->subdomain);  // poisson.cpp:228-229
    else
        s->domain->generalized_minimum_residual(s->u, s->f, *s->subdomain); // poisson.cpp:230-231
    if (num_iterations) *num_iterations = s->domain->num_iterations;
    int n = (int)s->domain->history.size();
    if (history_len) *history_len = n;
    if (history)
        for (int i = 0; i < n && i < history_cap; i++) history[i] = s->domain->history[i];
    return 0;
}

int prfdd_solver_solve(prfdd_solver *s, int solver_id, int *num_iterations, double *history, int history_cap, int *history_len)
{
    return guarded(s, [&]() {
        int rc = do_solve(s, solver_id, num_iterations, history, history_cap, history_len);
        device.finish();
        return rc;
    });
}

int prfdd_solver_solve_host(prfdd_solver *s, int solver_id, const double *f_host, double *u_host, int *num_iterations, double *history, int history_cap, int *history_len)
{
    return guarded(s, [&]() {
        const size_t bytes = sizeof(double) * (size_t)s->domain->num_local_points;
        // page-locked caller buffers are used directly; pageable ones are staged through the solver's pinned buffers
        auto pinned = [](const void *p) {
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
            return a.type == cudaMemoryTypeHost;
        };
        const void *src = f_host;
        if (!pinned(f_host)) { memcpy(s->pin_in, f_host, bytes); src = s->pin_in; }
        dev::check(cudaMemcpyAsync(s->f.ptr(), src, bytes, cudaMemcpyHostToDevice, s->stream), "solve_host/h2d");
        int rc = do_solve(s, solver_id, num_iterations, history, history_cap, history_len);
        const bool out_pinned = pinned(u_host);
        dev::check(cudaMemcpyAsync(out_pinned ? (void *)u_host : (void *)s->pin_out, s->u.ptr(), bytes, cudaMemcpyDeviceToHost, s->stream), "solve_host/d2h");
        device.finish();
        if (!out_pinned) memcpy(u_host, s->pin_out, bytes);
        return rc;
    });
}

long long prfdd_solver_query(prfdd_solver *s, int what)
{
    Domain<STYPE> *d = s->domain;
    Subdomain<PTYPE> *sd = s->subdomain.get();
    switch (what)
    {
    case PRFDD_Q_DIM: return s->dim_;
    case PRFDD_Q_NUM_LOCAL_ELEMENTS: return d->num_local_elements;
    case PRFDD_Q_NUM_LOCAL_POINTS: return d->num_local_points;
    case PRFDD_Q_NUM_LOCAL_NODES: return d->num_local_nodes;
    case PRFDD_Q_NUM_BDARY_NODES: return d->num_boundary_nodes();
    case PRFDD_Q_NUM_TOTAL_ELEMENTS: return d->num_total_elements;
    case PRFDD_Q_NUM_GLOBAL_NODES: return d->num_total_nodes;
    default: return sd ? sd->query(what) : -1;
    }
}

long long prfdd_solver_get_array(prfdd_solver *s, int what, void *dst, long long cap)
{
    long long result = -1;
    guarded(s, [&]() {
        Domain<STYPE> *d = s->domain;
        auto put = [&](const void *src, size_t elem, long long count) {
            if ((long long)(elem * count) > cap) { result = -(long long)(elem * count); return; }
            memcpy(dst, src, elem * count);
            result = count;
        };
        auto put_dev = [&](const dev::memory &m, size_t elem, long long count) {
            if ((long long)(elem * count) > cap) { result = -(long long)(elem * count); return; }
            m.copyTo(dst, elem * count);
            result = count;
        };
        switch (what)
        {
        case PRFDD_A_NODE_OF_POINT: put(d->Q_matrix().col_hst.data(), sizeof(int), d->num_local_points); break;
        case PRFDD_A_BOUNDARY_NODES: put(d->boundary_node_ids().data(), sizeof(long long), d->num_boundary_nodes()); break;
        case PRFDD_A_ASSEMBLED_WEIGHT: put_dev(d->assembled_weight_dev(), sizeof(double), d->num_local_nodes); break;
        case PRFDD_A_D_HAT: put(d->D_hat_hst.data(), sizeof(double), (long long)d->D_hat_hst.size()); break;
        case PRFDD_A_U: put_dev(s->u, sizeof(double), d->num_local_points); break;
        case PRFDD_A_U_STAR: put_dev(s->u_star, sizeof(double), d->num_local_points); break;
        case PRFDD_A_F: put_dev(s->f, sizeof(double), d->num_local_points); break;
        default:
            if (s->subdomain) result = s->subdomain->get_array(what, dst, cap);
        }
        return 0;
    });
    return result;
}

int prfdd_solver_apply(prfdd_solver *s, int what, const double *in_host, double *out_host)
{
    return guarded(s, [&]() {
        Domain<STYPE> *d = s->domain;
        const int P = d->num_local_points;
        if (what == PRFDD_APPLY_STIFFNESS || what == PRFDD_APPLY_DSSUM || what == PRFDD_APPLY_DSSUM_WEIGHTED || what == PRFDD_APPLY_PRECONDITIONER)
        {
            if (!s->apply_in.is_initialized())
            {
                s->apply_in = device.malloc<double>(P);
                s->apply_out = device.malloc<double>(P);
            }
            dev::memory &a = s->apply_in, &b = s->apply_out;
            a.copyFrom(in_host, sizeof(double) * P);
            if (what == PRFDD_APPLY_STIFFNESS) d->stiffness_matrix(b, a);
            else if (what == PRFDD_APPLY_DSSUM) d->direct_stiffness_summation(b, a, true, false);
            else if (what == PRFDD_APPLY_DSSUM_WEIGHTED) d->direct_stiffness_summation(b, a, true, true);
            else
            {
                if (!s->opt.use_preconditioner) throw std::runtime_error("preconditioner is off");
                if (s->opt.preconditioner_type == 0) s->subdomain->flexible_conjugate_gradient(b, a);
                else s->subdomain->generalized_minimum_residual(b, a);
            }
            b.copyTo(out_host, sizeof(double) * P);
            return 0;
        }
        if (!s->subdomain) return -1;
        return s->subdomain->apply(what, in_host, out_host);
    });
}

int prfdd_solver_profile_vcycle(prfdd_solver *s, int reps, char *text, int capacity)
{
    return guarded(s, [&]() { return s->subdomain ? s->subdomain->profile_vcycle(reps, text, capacity) : -1; });
}

int prfdd_solver_time_spmv(prfdd_solver *s, int reps, double out[6])
{
    return guarded(s, [&]() { return s->subdomain ? s->subdomain->time_spmv(reps, out) : -1; });
}

// ---- host-only AMG setup handle ------------------------------------------------------------
struct prfdd_amg_host
{
    amg::Hierarchy H;
};

int prfdd_amg_host_setup(prfdd_amg_host **h, int n, const int *ptr, const int *col, const double *val, int cheby_order, int max_coarse_size)
{
    return prfdd_amg_host_setup_ex(h, n, ptr, col, val, cheby_order, max_coarse_size, -1);
}

int prfdd_amg_host_setup_ex(prfdd_amg_host **h, int n, const int *ptr, const int *col, const double *val, int cheby_order, int max_coarse_size, int coarsening)
{
    try
    {
        amg::HostCSR A;
        A.num_rows = A.num_cols = n;
        A.ptr.assign(ptr, ptr + n + 1);
        A.col.assign(col, col + ptr[n]);
        A.val.assign(val, val + ptr[n]);
        prfdd_amg_host *o = new prfdd_amg_host();
        o->H.coarsening = coarsening;
        o->H.setup(std::move(A), cheby_order, max_coarse_size, 0.25, 4, 25, /*on_device=*/false);
        *h = o;
        return 0;
    }
    catch (const std::exception &ex)
    {
        fprintf(stderr, "prfdd_amg_host_setup: %s\n", ex.what());
        return -1;
    }
}

int prfdd_amg_host_destroy(prfdd_amg_host *h) { delete h; return 0; }
int prfdd_amg_host_num_levels(const prfdd_amg_host *h) { return h->H.num_levels(); }

int prfdd_amg_host_level_sizes(const prfdd_amg_host *h, int level, int sizes[4])
{
    const amg::Level &L = h->H.levels[level];
    sizes[0] = L.n; sizes[1] = L.A.nnz(); sizes[2] = L.P.num_cols; sizes[3] = L.P.nnz();
    return 0;
}

int prfdd_amg_host_get_matrix(const prfdd_amg_host *h, int level, int which, int *ptr, int *col, double *val)
{
    const amg::Level &L = h->H.levels[level];
    const amg::HostCSR &M = which == 0 ? L.A : L.P;
    if (M.ptr.empty()) return -1;
    std::copy(M.ptr.begin(), M.ptr.end(), ptr);
    std::copy(M.col.begin(), M.col.end(), col);
    std::copy(M.val.begin(), M.val.end(), val);
    return 0;
}

int prfdd_amg_host_get_vectors(const prfdd_amg_host *h, int level, signed char *cf, double *ds, double *coefs, double eigs[2])
{
    const amg::Level &L = h->H.levels[level];
    if (cf) std::copy(L.cf.begin(), L.cf.end(), cf);
    if (ds) std::copy(L.ds_hst.begin(), L.ds_hst.end(), ds);
    if (coefs) std::copy(L.coefs.begin(), L.coefs.end(), coefs);
    if (eigs) { eigs[0] = L.max_eig; eigs[1] = L.min_eig; }
    return 0;
}

int prfdd_amg_host_get_coarse_inverse(const prfdd_amg_host *h, double *Ainv)
{
    std::copy(h->H.Ainv_hst.begin(), h->H.Ainv_hst.end(), Ainv);
    return 0;
}

int prfdd_solver_output(prfdd_solver *s, const char *output_name)
{
    // poisson.cpp:233-235: domain.output("domain", 3, "u_star", u_star, "f", f, "u", u); one piece per rank here
    return guarded(s, [&]() {
        Domain<STYPE> *d = s->domain;
        const long long P = d->num_local_points;
        std::vector<double> x(P), y(P), z(P), us(P), ff(P), uu(P);
        long long at = 0;
        for (auto &e : d->elements)
            for (size_t v = 0; v < e.x.size(); v++, at++)
            {
                x[at] = e.x[v];
                y[at] = e.y.empty() ? 0.0 : e.y[v];
                z[at] = e.z.empty() ? 0.0 : e.z[v];
            }
        s->u_star.copyTo(us.data(), P * sizeof(double));
        s->f.copyTo(ff.data(), P * sizeof(double));
        s->u.copyTo(uu.data(), P * sizeof(double));
        char path[1024];
        snprintf(path, sizeof(path), "%s_%d.vtk", output_name, prfdd_host::proc_id);
        const char *names[3] = {"u_star", "f", "u"};
        const double *fields[3] = {us.data(), ff.data(), uu.data()};
        return prfdd_write_vtk(path, prfdd_host::dim, d->poly_degree + 1, d->num_local_elements, x.data(), y.data(), z.data(), 3, names, fields);
    });
}

int prfdd_solver_output_subdomain(prfdd_solver *s, const char *output_name)
{
    return guarded(s, [&]() {
        if (!s->subdomain || !s->opt.use_preconditioner) return -8;
        return s->subdomain->output_last_application(output_name);
    });
}

double prfdd_solver_timer_total(prfdd_solver *s, const char *key)
{
    if (!strcmp(key, "__enable__")) { s->tmr.enabled = true; return 0.0; }
    if (!strcmp(key, "__disable__")) { s->tmr.enabled = false; return 0.0; }
    return s->tmr.total(key);
}

} // extern "C"
