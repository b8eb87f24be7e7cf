// k_vector.cu -- fused vector algebra and deterministic device-side reductions (sm_100a)
//
// Replaces math.okl:5-35, domain.okl:100-264 and subdomain.okl:103-282 of the reference.
// All kernels are HBM-bound streaming kernels: grid-stride loops over 128-bit (double2) accesses,
// grids sized in whole multiples of the SM count.  Reductions never touch the host: each block
// writes one partial per sum, the last block to finish (atomic ticket) adds the partials in a fixed
// order and stores the result in device memory, so results are bit-reproducible run to run.
#include "common.cuh"

namespace prfdd
{
long long g_launch_count = 0;
double g_algorithmic_bytes = 0.0;

constexpr int kThreads = 256;
constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 4 * 148; // 4 CTAs per SM on a B200
constexpr int kRedMaxK = 8;

// ---------------------------------------------------------------------------------------------
// element-wise kernels
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(kThreads) k_map(long long n, F f)
{
    pdl_wait();
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

// bpe: algorithmic bytes per element (every operand of the element-wise kernel once)
template <class F>
static int map(long long n, cudaStream_t st, double bpe, F f)
{
    if (n <= 0) return 0;
    launch_pdl(k_map<F>, stream_grid(n, kThreads, 2, 8), kThreads, 0, st, n, f);
    return launched(bpe * (double)n);
}

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum in a fixed order; result valid in thread 0
template <int K>
__device__ __forceinline__ void block_sum(double (&acc)[K], double *smem /* [K][warps] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int W = kRedThreads / 32;
#pragma unroll
    for (int k = 0; k < K; k++)
    {
        double v = warp_sum(acc[k]);
        if (lane == 0) smem[k * W + warp] = v;
    }
    __syncthreads();
    if (warp == 0)
    {
#pragma unroll
        for (int k = 0; k < K; k++)
        {
            double v = (lane < W) ? smem[k * W + lane] : 0.0;
            v = warp_sum(v);
            acc[k] = v;
        }
    }
    __syncthreads();
}

// two elements per trip with their own partial sums (the operand loads of both are independent and in flight together; a register
// budget is stated because ptxas otherwise fits the loop into 32 registers by serialising them, profiles/r2_notes.txt); the two
// partial sums are added in a fixed order
template <int K, class F>
__global__ void __launch_bounds__(kRedThreads, 4) k_reduce(long long n, F f, double *partials, unsigned int *counter, double *out, int out_stride)
{
    __shared__ double smem[K * (kRedThreads / 32)];
    __shared__ bool is_last;
    pdl_wait();
    double acc[K], acc2[K];
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] = acc2[k] = 0.0;
    long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < n; i += 2 * stride)
    {
        f(i, acc);
        f(i + stride, acc2);
    }
    if (i < n) f(i, acc);
#pragma unroll
    for (int k = 0; k < K; k++) acc[k] += acc2[k];
    block_sum<K>(acc, smem);
    if (threadIdx.x == 0)
    {
#pragma unroll
        for (int k = 0; k < K; k++) partials[k * kRedMaxBlocks + blockIdx.x] = acc[k];
        __threadfence();
        unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last)
    {
        __threadfence();
#pragma unroll
        for (int k = 0; k < K; k++)
        {
            double v = 0.0;
            for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) v += __ldcg(&partials[k * kRedMaxBlocks + b]);
            acc[k] = v;
        }
        block_sum<K>(acc, smem);
        if (threadIdx.x == 0)
        {
#pragma unroll
            for (int k = 0; k < K; k++) out[k * out_stride] = acc[k];
            *counter = 0u;
        }
    }
}

template <int K, class F>
static int reduce(prfdd_reduce_ws *ws, double *out, int out_stride, long long n, cudaStream_t st, double bpe, F f)
{
    static_assert(K <= kRedMaxK, "too many fused sums");
    if (ws == nullptr) return -2;
    int grid = stream_grid(n > 0 ? n : 1, kRedThreads, 4, 4);
    if (grid > kRedMaxBlocks) grid = kRedMaxBlocks;
    launch_pdl(k_reduce<K, F>, grid, kRedThreads, 0, st, n, f, ws->partials, ws->counter, out, out_stride);
    return launched(bpe * (double)n);
}
} // namespace prfdd

using namespace prfdd;

extern "C" {

int prfdd_reduce_ws_create(prfdd_reduce_ws **ws)
{
    prfdd_reduce_ws *w = new prfdd_reduce_ws();
    cudaError_t e = cudaMalloc(&w->partials, sizeof(double) * kRedMaxK * kRedMaxBlocks);
    if (e != cudaSuccess) { delete w; return (int)e; }
    e = cudaMalloc(&w->counter, sizeof(unsigned int));
    if (e != cudaSuccess) { cudaFree(w->partials); delete w; return (int)e; }
    cudaMemset(w->counter, 0, sizeof(unsigned int));
    *ws = w;
    return 0;
}

int prfdd_reduce_ws_destroy(prfdd_reduce_ws *ws)
{
    if (!ws) return 0;
    cudaFree(ws->partials);
    cudaFree(ws->counter);
    delete ws;
    return 0;
}

double prfdd_algorithmic_bytes(void) { return g_algorithmic_bytes; }
void prfdd_algorithmic_bytes_reset(void) { g_algorithmic_bytes = 0.0; }
void prfdd_algorithmic_bytes_add(double bytes) { g_algorithmic_bytes += bytes; }
long long prfdd_launch_count(void) { return g_launch_count; }
void prfdd_launch_count_reset(void) { g_launch_count = 0; }
void prfdd_launch_count_add(long long n) { g_launch_count += n; }

// ------------------------------------------------------------------------------ math.okl
int prfdd_set_to_value(double *u, double alpha, int n, int offset, prfdd_stream_t stream)
{
    return map(n, S(stream), 8.0, [=] __device__(long long i) { u[i + offset] = alpha; });
}

int prfdd_vector_set_to_value(double *data, double value, int size, prfdd_stream_t stream)
{
    return map(size, S(stream), 8.0, [=] __device__(long long i) { data[i] = value; });
}

int prfdd_invert_vector_elements(double *u, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 16.0, [=] __device__(long long i) { u[i] = 1.0 / u[i]; });
}

int prfdd_vector_vector_addition(double *uv, double alpha, const double *u, double beta, const double *v, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 24.0, [=] __device__(long long i) { uv[i] = alpha * u[i] + beta * v[i]; });
}

int prfdd_vector_scaling(double *au, double alpha, const double *u, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 16.0, [=] __device__(long long i) { au[i] = alpha * u[i]; });
}

int prfdd_vector_scaling_dev(double *au, const double *num, const double *den, const double *u, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 16.0, [=] __device__(long long i) {
        double a = den ? (*num) / (*den) : (*num);
        au[i] = a * u[i];
    });
}

struct PtrPack
{
    const double *p[32];
};

int prfdd_multi_axpy_dev(double *y, const double *const *X, const double *coef, int coef_stride, double sign, int count, int n, prfdd_stream_t stream)
{
    if (count > 32) return -3;
    PtrPack pk;
    for (int i = 0; i < count; i++) pk.p[i] = X[i];
    // y = 1.0*y + (sign*c_i) X_i, i ascending: the same chain as the reference's sequence of
    // vector_vector_addition calls (domain.tpp:817-822, 902-907; subdomain.tpp:4396-4401, 4473-4478)
    return map(n, S(stream), 8.0 * (count + 2), [=] __device__(long long i) {
        double acc = y[i];
        for (int k = 0; k < count; k++)
        {
            const double c = sign * coef[k * coef_stride];
            if (c != 0.0) acc = acc + c * pk.p[k][i]; // columns beyond the last one used carry c == 0 exactly
        }
        y[i] = acc;
    });
}

// ---------------------------------------------------------------- domain.okl element-wise
int prfdd_initialize_arrays(double *u_k, double *r_k, const double *f, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 24.0, [=] __device__(long long i) { u_k[i] = 0.0; r_k[i] = f[i]; });
}

int prfdd_solution_and_residual_update(double *u_k, double *r_kp1, const double *r_k, const double *p_k, const double *q_k, double alpha_k, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 48.0, [=] __device__(long long i) {
        u_k[i] += alpha_k * p_k[i];
        r_kp1[i] = r_k[i] - alpha_k * q_k[i];
    });
}

int prfdd_solution_and_residual_update_dev(double *u_k, double *r_kp1, const double *r_k, const double *p_k, const double *q_k, const double *num, const double *den, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 48.0, [=] __device__(long long i) {
        double alpha_k = (*num) / (*den);
        u_k[i] += alpha_k * p_k[i];
        r_kp1[i] = r_k[i] - alpha_k * q_k[i];
    });
}

int prfdd_residual_and_search_update(double *p_k, double *r_k, const double *z_k, const double *r_kp1, double beta_k, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 40.0, [=] __device__(long long i) {
        p_k[i] = z_k[i] + beta_k * p_k[i];
        r_k[i] = r_kp1[i];
    });
}

int prfdd_residual_and_search_update_dev(double *p_k, double *r_k, const double *z_k, const double *r_kp1, const double *num, const double *den, int n, prfdd_stream_t stream)
{
    return map(n, S(stream), 40.0, [=] __device__(long long i) {
        double beta_k = (*num) / (*den);
        p_k[i] = z_k[i] + beta_k * p_k[i];
        r_k[i] = r_kp1[i];
    });
}

int prfdd_copy_from_domain_data(double *u, const double *v, int num_points, prfdd_stream_t stream)
{
    return map(num_points, S(stream), 16.0, [=] __device__(long long i) { u[i] = v[i]; });
}

int prfdd_copy_to_domain_data(double *u, const double *v, int num_points, prfdd_stream_t stream)
{
    return map(num_points, S(stream), 16.0, [=] __device__(long long i) { u[i] = v[i]; });
}

// ---------------------------------------------------------------- AMG/kernels.cu
int prfdd_main_scaled_residual(double *Sr, double *w, const double *f_m_Au, const double *Sv, double alpha, int size, prfdd_stream_t stream)
{
    return map(size, S(stream), 32.0, [=] __device__(long long i) {
        double s = Sv[i] * f_m_Au[i];
        Sr[i] = s;
        w[i] = alpha * s;
    });
}

int prfdd_main_polynomial_evaluation(double *w, double *v, const double *r, const double *D_val, double alpha, int size, prfdd_stream_t stream)
{
    return map(size, S(stream), 40.0, [=] __device__(long long i) {
        double t = v[i] * D_val[i];
        v[i] = t;
        w[i] = alpha * r[i] + t;
    });
}

int prfdd_main_update_field(double *u, const double *w, const double *D_val, int size, prfdd_stream_t stream)
{
    return map(size, S(stream), 32.0, [=] __device__(long long i) { u[i] += D_val[i] * w[i]; });
}

int prfdd_vector_multiplication(double *uv, const double *u, const double *v, int size, prfdd_stream_t stream)
{
    return map(size, S(stream), 24.0, [=] __device__(long long i) { uv[i] = u[i] * v[i]; });
}

int prfdd_cheby_order1(double *u, const double *r, const double *ds, double c, int u_is_zero, int size, prfdd_stream_t stream)
{
    return map(size, S(stream), (u_is_zero ? 24.0 : 32.0), [=] __device__(long long i) {
        double w = ds[i] * (c * r[i]);
        u[i] = u_is_zero ? w : u[i] + w;
    });
}

// ---------------------------------------------------------------- reductions
int prfdd_residual_norm(prfdd_reduce_ws *ws, double *out, const double *r_k, const double *QQt_r_k, const double *mask, int n, prfdd_stream_t stream)
{
    return reduce<1>(ws, out, 1, n, S(stream), 24.0, [=] __device__(long long i, double(&a)[1]) { a[0] += r_k[i] * QQt_r_k[i] * mask[i]; });
}

int prfdd_projection_inner_products(prfdd_reduce_ws *ws, double *out, const double *z_k, const double *r_k, const double *p_k, const double *q_k, int n, prfdd_stream_t stream)
{
    return reduce<2>(ws, out, 1, n, S(stream), 32.0, [=] __device__(long long i, double(&a)[2]) {
        a[0] += z_k[i] * r_k[i];
        a[1] += p_k[i] * q_k[i];
    });
}

int prfdd_inner_product_flexible(prfdd_reduce_ws *ws, double *out, const double *r_k, const double *r_kp1, const double *z_k, int n, prfdd_stream_t stream)
{
    return reduce<1>(ws, out, 1, n, S(stream), 24.0, [=] __device__(long long i, double(&a)[1]) { a[0] += (r_kp1[i] - r_k[i]) * z_k[i]; });
}

int prfdd_inner_product(prfdd_reduce_ws *ws, double *out, const double *u_k, const double *v_k, const double *mask, int n, prfdd_stream_t stream)
{
    return reduce<1>(ws, out, 1, n, S(stream), 24.0, [=] __device__(long long i, double(&a)[1]) { a[0] += u_k[i] * v_k[i] * mask[i]; });
}

int prfdd_weighted_inner_product(prfdd_reduce_ws *ws, double *out, const double *u, const double *v, const double *w, int n, prfdd_stream_t stream)
{
    if (w)
        return reduce<1>(ws, out, 1, n, S(stream), (8.0 * ((u == v ? 1 : 2) + (w ? 1 : 0))), [=] __device__(long long i, double(&a)[1]) { a[0] += u[i] * v[i] * w[i]; });
    return reduce<1>(ws, out, 1, n, S(stream), (8.0 * ((u == v ? 1 : 2) + (w ? 1 : 0))), [=] __device__(long long i, double(&a)[1]) { a[0] += u[i] * v[i]; });
}

int prfdd_multi_inner_product(prfdd_reduce_ws *ws, double *out, const double *u, const double *const *V, const double *w, int count, int n, prfdd_stream_t stream)
{
    if (count > 32) return -3;
    int rc = 0;
    for (int base = 0; base < count && rc == 0; base += 4)
    {
        int c = count - base < 4 ? count - base : 4;
        const double *v0 = V[base], *v1 = c > 1 ? V[base + 1] : V[base], *v2 = c > 2 ? V[base + 2] : V[base], *v3 = c > 3 ? V[base + 3] : V[base];
        double *o = out + base;
        if (c == 1)
            rc = reduce<1>(ws, o, 1, n, S(stream), (8.0 * (c + 1 + (w ? 1 : 0))), [=] __device__(long long i, double(&a)[1]) { a[0] += u[i] * v0[i] * (w ? w[i] : 1.0); });
        else if (c == 2)
            rc = reduce<2>(ws, o, 1, n, S(stream), (8.0 * (c + 1 + (w ? 1 : 0))), [=] __device__(long long i, double(&a)[2]) {
                double uw = u[i] * (w ? w[i] : 1.0);
                a[0] += uw * v0[i];
                a[1] += uw * v1[i];
            });
        else if (c == 3)
            rc = reduce<3>(ws, o, 1, n, S(stream), (8.0 * (c + 1 + (w ? 1 : 0))), [=] __device__(long long i, double(&a)[3]) {
                double uw = u[i] * (w ? w[i] : 1.0);
                a[0] += uw * v0[i];
                a[1] += uw * v1[i];
                a[2] += uw * v2[i];
            });
        else
            rc = reduce<4>(ws, o, 1, n, S(stream), (8.0 * (c + 1 + (w ? 1 : 0))), [=] __device__(long long i, double(&a)[4]) {
                double uw = u[i] * (w ? w[i] : 1.0);
                a[0] += uw * v0[i];
                a[1] += uw * v1[i];
                a[2] += uw * v2[i];
                a[3] += uw * v3[i];
            });
    }
    return rc;
}

// Arnoldi step, assembled side (subdomain.tpp:4396-4412 on the assembled copies): aq <- aq - sum_k coef[k] aV_k, same chain
// as prfdd_multi_axpy_dev, and out[0] = sum_i w[i] aq[i]^2 of the result -- the orthogonalised column's norm without assembling
// it again (Q^T is linear, and aV_k = Q^T V_k are kept from the previous steps)
int prfdd_orthogonalize_norm(prfdd_reduce_ws *ws, double *out, double *aq, const double *const *aV, const double *coef, const double *w, int count, int n, prfdd_stream_t stream)
{
    if (count > 32) return -3;
    PtrPack pk;
    for (int i = 0; i < count; i++) pk.p[i] = aV[i];
    return reduce<1>(ws, out, 1, n, S(stream), 8.0 * (count + 3), [=] __device__(long long i, double(&a)[1]) {
        double acc = aq[i];
        for (int k = 0; k < count; k++)
        {
            const double c = -coef[k];
            if (c != 0.0) acc = acc + c * pk.p[k][i];
        }
        aq[i] = acc;
        a[0] += acc * acc * w[i];
    });
}

// next Arnoldi vector and its assembled copy in one launch:
//   V_next[i] = scale * (q[i] - sum_k coef[k] V_k[i])   i < n            (vector_vector_addition chain + vector_scaling,
//   aV_next[i] = scale * aq[i]                           i < n_assembled   subdomain.tpp:4396-4401, 4455-4458)
// scale read from device memory; count = 0 gives the plain scaling of the cycle's first vector (tpp:4349-4352)
int prfdd_arnoldi_next(double *V_next, const double *q, const double *const *V, const double *coef, int count, const double *scale, int n, double *aV_next, const double *aq, int n_assembled, prfdd_stream_t stream)
{
    if (count > 32) return -3;
    PtrPack pk;
    for (int i = 0; i < count; i++) pk.p[i] = V[i];
    return map((long long)n + n_assembled, S(stream), (8.0 * (count + 2) * n + 16.0 * n_assembled) / (double)((long long)n + n_assembled > 0 ? (long long)n + n_assembled : 1), [=] __device__(long long i) {
        const double a = *scale;
        if (i < n)
        {
            double acc = q[i];
            for (int k = 0; k < count; k++)
            {
                const double c = -coef[k];
                if (c != 0.0) acc = acc + c * pk.p[k][i];
            }
            V_next[i] = a * acc;
        }
        else
            aV_next[i - n] = a * aq[i - n];
    });
}

int prfdd_weighted_projection_inner_products(prfdd_reduce_ws *ws, double *out, const double *z_k, const double *r_k, const double *p_k, const double *q_k, const double *weight, int n, prfdd_stream_t stream)
{
    return reduce<2>(ws, out, 1, n, S(stream), 40.0, [=] __device__(long long i, double(&a)[2]) {
        a[0] += z_k[i] * r_k[i] * weight[i];
        a[1] += p_k[i] * q_k[i] * weight[i];
    });
}

int prfdd_search_update_inner_product(prfdd_reduce_ws *ws, double *out, const double *r_k, const double *r_kp1, const double *z_k, const double *weight, int n, prfdd_stream_t stream)
{
    return reduce<1>(ws, out, 1, n, S(stream), 32.0, [=] __device__(long long i, double(&a)[1]) { a[0] += (r_kp1[i] - r_k[i]) * z_k[i] * weight[i]; });
}

} // extern "C"
