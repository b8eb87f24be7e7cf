// k_krylov.cu -- one-thread kernels that keep the inner Krylov solvers' scalar bookkeeping on the device
// (Hessenberg column, Givens rotations, stop tests, back-substitution; FCG alpha/beta).
// Restates the host code of subdomain.tpp:4341-4478 (GMRES) and 4203-4261 (FCG) so that a whole
// preconditioner application needs no host synchronisation.
#include "common.cuh"

namespace prfdd
{
constexpr int MV = PRFDD_KRYLOV_MAXV;

// Outer flexible CG without host round trips (Domain::flexible_conjugate_gradient, domain.tpp:611-725, as one CUDA graph whose loop
// is a conditional WHILE node): this kernel is the loop test of domain.tpp:683-695.  state[0] = number of residual norms recorded so
// far, state[1] = number of completed search-direction updates; hist[k] = k-th residual norm; hist[-1..] untouched.  It records
// sqrt(*sum), decides exactly as the host loop does (same IEEE sqrt and division) and sets the graph's condition.
__global__ void k_fcg_outer_check(const double *sum, double *hist, int *state, double tolerance, int use_relative, int max_iterations, unsigned long long handle, int has_handle)
{
    const double r_norm = sqrt(*sum);
    const int k = state[0];
    hist[k] = r_norm;
    state[0] = k + 1;
    unsigned int go = 1;
    if (k >= 1)
    {
        // k-th norm = the residual after the first half of iteration k - 1
        const bool converged = use_relative ? (r_norm / hist[0] < tolerance) : (r_norm < tolerance);
        if (converged || isnan(r_norm) || k >= max_iterations) go = 0;
    }
    if (has_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)handle, go);
}

__global__ void k_fcg_outer_count(int *state) { state[1] += 1; }
__global__ void k_fcg_outer_reset(int *state) { state[0] = 0; state[1] = 0; }

__global__ void k_krylov_reset(prfdd_krylov_state *st)
{
    st->stopped = 0;
    st->cycle_active = 1;
    st->j_last = -1;
    st->alpha_cg = 0.0;
    st->beta_cg = 0.0;
    st->one = 1.0;
    for (int i = 0; i < MV; i++) st->y[i] = 0.0;
}

__global__ void k_gmres_begin_cycle(prfdd_krylov_state *st, int first_cycle)
{
    const double g0 = sqrt(st->red[0]);
    st->gamma[0] = g0;
    if (first_cycle) st->r0_norm = g0;
    st->inv_gamma0 = 1.0 / g0;
    st->cycle_active = st->stopped ? 0 : 1;
    st->j_last = -1;
}

__global__ void k_gmres_column(prfdd_krylov_state *st, int j, int iter, int max_iterations, double tolerance, int use_relative)
{
    if (st->stopped) return;
    double *H = st->H;
    for (int i = 0; i <= j; i++) H[i * MV + j] = st->hcol[i];
    // Givens rotations on the new column (subdomain.tpp:4404-4409); H[j+1][j] is not stored (it is alpha_j)
    for (int i = 0; i < j; i++)
    {
        const double h_ij = H[i * MV + j];
        H[i * MV + j] = st->c[i] * h_ij + st->s[i] * H[(i + 1) * MV + j];
        H[(i + 1) * MV + j] = -st->s[i] * h_ij + st->c[i] * H[(i + 1) * MV + j];
    }
    const double alpha_j = sqrt(st->red[0]);
    st->iterations += 1;
    if (fabs(alpha_j) == 0.0)
    {
        // tpp:4415-4419: break BEFORE this column is rotated into the solution: columns 0..j are still used
        st->stopped = 1;
        st->j_last = j;
        st->inv_alpha = 0.0;
        return;
    }
    const double beta_j = sqrt(H[j * MV + j] * H[j * MV + j] + alpha_j * alpha_j);
    const double gamma_j = 1.0 / beta_j;
    st->c[j] = H[j * MV + j] * gamma_j;
    st->s[j] = alpha_j * gamma_j;
    H[j * MV + j] = beta_j;
    st->gamma[j + 1] = -st->s[j] * st->gamma[j];
    st->gamma[j] = st->c[j] * st->gamma[j];
    const double r_norm = fabs(st->gamma[j + 1]);
    st->r_norm = r_norm;
    st->j_last = j;
    st->inv_alpha = 1.0 / alpha_j;
    bool stop = use_relative ? (r_norm / st->r0_norm < tolerance) : (r_norm < tolerance);
    if (iter >= max_iterations) stop = true;
    if (stop) st->stopped = 1;
}

__global__ void k_gmres_end_cycle(prfdd_krylov_state *st, int num_vectors)
{
    for (int i = 0; i < MV; i++) st->y[i] = 0.0;
    if (!st->cycle_active) return;
    int j = st->j_last;
    if (j < 0) return;
    if (j >= num_vectors) j = num_vectors - 1;
    // note: when the loop broke on alpha_j == 0 the reference back-substitutes with the un-updated c/s of
    // column j (tpp:4460-4470 runs with whatever H[j][j] holds); the same happens here
    double *H = st->H;
    double cc[MV];
    for (int k = j; k >= 0; k--)
    {
        double gamma_k = st->gamma[k];
        for (int i = j; i > k; i--) gamma_k -= H[k * MV + i] * cc[i];
        cc[k] = gamma_k / H[k * MV + k];
    }
    for (int i = 0; i <= j; i++) st->y[i] = cc[i];
}

__global__ void k_fcg_alpha(prfdd_krylov_state *st)
{
    if (st->stopped) { st->alpha_cg = 0.0; return; }
    st->gamma_cg = st->red[0];
    st->alpha_cg = st->red[0] / st->red[1];
}

__global__ void k_fcg_check(prfdd_krylov_state *st, int iter, int max_iterations, double tolerance, int use_relative)
{
    if (st->stopped) return;
    const double r_norm = sqrt(st->red[2]);
    st->r_norm = r_norm;
    st->iterations += 1;
    bool stop = use_relative ? (r_norm / st->r0_norm < tolerance) : (r_norm < tolerance);
    if (iter == max_iterations) stop = true;
    if (stop) st->stopped = 1;
}

__global__ void k_fcg_beta(prfdd_krylov_state *st)
{
    if (st->stopped) { st->beta_cg = 0.0; return; }
    st->beta_cg = st->red[3] / st->gamma_cg;
}

__global__ void __launch_bounds__(256) k_axpy_dev(double *__restrict__ y, const double *__restrict__ a, double sign, const double *__restrict__ x, long long n)
{
    const double c = sign * (*a);
    if (c == 0.0) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] += c * x[i];
}

__global__ void __launch_bounds__(256) k_search_update_gated(double *__restrict__ p, double *__restrict__ r, const double *__restrict__ z, const double *__restrict__ r1, const double *__restrict__ beta, const int *__restrict__ skip, long long n)
{
    if (*skip) return;
    const double b = *beta;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
        p[i] = z[i] + b * p[i];
        r[i] = r1[i];
    }
}
} // namespace prfdd

using namespace prfdd;

extern "C" {

int prfdd_krylov_reset(prfdd_krylov_state *st, prfdd_stream_t stream)
{
    k_krylov_reset<<<1, 1, 0, S(stream)>>>(st);
    return launched();
}

int prfdd_fcg_outer_reset(int *state, prfdd_stream_t stream)
{
    k_fcg_outer_reset<<<1, 1, 0, S(stream)>>>(state);
    return launched(0.0);
}
int prfdd_fcg_outer_check(const double *sum, double *hist, int *state, double tolerance, int use_relative, int max_iterations, unsigned long long cond_handle, int has_handle, prfdd_stream_t stream)
{
    k_fcg_outer_check<<<1, 1, 0, S(stream)>>>(sum, hist, state, tolerance, use_relative, max_iterations, cond_handle, has_handle);
    return launched(0.0);
}
int prfdd_fcg_outer_count(int *state, prfdd_stream_t stream)
{
    k_fcg_outer_count<<<1, 1, 0, S(stream)>>>(state);
    return launched(0.0);
}
int prfdd_gmres_begin_cycle(prfdd_krylov_state *st, int first_cycle, prfdd_stream_t stream)
{
    k_gmres_begin_cycle<<<1, 1, 0, S(stream)>>>(st, first_cycle);
    return launched();
}
int prfdd_gmres_column(prfdd_krylov_state *st, int j, int iter, int max_iterations, double tolerance, int use_relative, prfdd_stream_t stream)
{
    if (j >= MV) return -3;
    k_gmres_column<<<1, 1, 0, S(stream)>>>(st, j, iter, max_iterations, tolerance, use_relative);
    return launched();
}
int prfdd_gmres_end_cycle(prfdd_krylov_state *st, int num_vectors, prfdd_stream_t stream)
{
    k_gmres_end_cycle<<<1, 1, 0, S(stream)>>>(st, num_vectors);
    return launched();
}
int prfdd_fcg_alpha(prfdd_krylov_state *st, prfdd_stream_t stream)
{
    k_fcg_alpha<<<1, 1, 0, S(stream)>>>(st);
    return launched();
}
int prfdd_fcg_check(prfdd_krylov_state *st, int iter, int max_iterations, double tolerance, int use_relative, prfdd_stream_t stream)
{
    k_fcg_check<<<1, 1, 0, S(stream)>>>(st, iter, max_iterations, tolerance, use_relative);
    return launched();
}
int prfdd_fcg_beta(prfdd_krylov_state *st, prfdd_stream_t stream)
{
    k_fcg_beta<<<1, 1, 0, S(stream)>>>(st);
    return launched();
}
int prfdd_axpy_dev(double *y, const double *a, double sign, const double *x, int n, prfdd_stream_t stream)
{
    if (n <= 0) return 0;
    k_axpy_dev<<<stream_grid(n, 256, 2, 8), 256, 0, S(stream)>>>(y, a, sign, x, n);
    return launched(24.0 * n);
}
int prfdd_residual_and_search_update_gated(double *p_k, double *r_k, const double *z_k, const double *r_kp1, const double *beta, const int *skip_flag, int n, prfdd_stream_t stream)
{
    if (n <= 0) return 0;
    k_search_update_gated<<<stream_grid(n, 256, 2, 8), 256, 0, S(stream)>>>(p_k, r_k, z_k, r_kp1, beta, skip_flag, n);
    return launched(40.0 * n);
}

} // extern "C"
