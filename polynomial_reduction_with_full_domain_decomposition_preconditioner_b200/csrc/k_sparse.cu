// k_sparse.cu -- gather-scatter, halo pack/unpack, CSR SpMV family, fused Chebyshev smoother (sm_100a)
//
// Replaces csr_matrix.okl:5-48 (one thread per row, scalar), the cusparseSpMV(CSR_ALG1) calls of
// AMG/csr_matrix.cpp:126-133 and the element-wise smoother kernels of AMG/kernels.cu:25-76.
//
// SpMV shape: a sub-warp of TPR lanes owns one row; lanes stride through the row so that col/val
// reads of neighbouring lanes are contiguous (coalesced), partial sums are combined with shuffles
// in a fixed order (deterministic).  Everything the reference does in separate element-wise launches
// around an SpMV (scaling by ds, c*r + v, u += ds*w, f - A u) is done in the SpMV epilogue, so one
// Chebyshev smoothing of order 2 is 2 launches and 2 passes over A instead of 7 launches.
#include "common.cuh"

namespace prfdd
{
constexpr int kSpThreads = 256;

template <int TPR>
__device__ __forceinline__ double row_dot(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ x, int row, int lane, bool valid)
{
    const int s = valid ? ptr[row] : 0, e = valid ? ptr[row + 1] : 0;
    double acc = 0.0;
    for (int j = s + lane; j < e; j += TPR) acc += val[j] * x[col[j]];
#pragma unroll
    for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, TPR);
    return acc;
}

// generic SpMV with an epilogue functor: epi(row, Ax) executed by lane 0 of the row's sub-warp.
// The row loop is warp-uniform (all 32 lanes take the same number of trips) so the full-mask
// shuffles are always executed by the whole warp.  RPG > 1: every sub-warp walks RPG rows at once
// (rows r, r + 32/TPR, ...), which multiplies the independent col/val -> x load chains a lane has in
// flight; the long rows of the coarse AMG levels are latency bound without it.
template <int TPR, int RPG, class Epi>
__global__ void __launch_bounds__(kSpThreads) k_spmv(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ x, int row_start, int num_rows, Epi epi)
{
    constexpr int RPW = 32 / TPR; // rows per warp and pass
    const int lane = threadIdx.x % TPR;
    const int sub = (threadIdx.x & 31) / TPR;
    const long long rows_per_grid = (long long)gridDim.x * (kSpThreads / TPR) * RPG;
    for (long long r0 = ((long long)blockIdx.x * (kSpThreads / 32) + (threadIdx.x >> 5)) * RPW * RPG; r0 < num_rows; r0 += rows_per_grid)
    {
        if constexpr (RPG == 1)
        {
            const long long r = r0 + sub;
            const bool valid = r < num_rows;
            const int row = row_start + (int)r;
            double ax = x ? row_dot<TPR>(ptr, col, val, x, row, lane, valid) : 0.0;
            if (valid && lane == 0) epi(row, ax);
        }
        else
        {
            int j[RPG], e[RPG];
            double acc[RPG];
#pragma unroll
            for (int g = 0; g < RPG; g++)
            {
                const long long r = r0 + g * RPW + sub;
                const bool valid = x && r < num_rows;
                j[g] = valid ? ptr[row_start + r] + lane : 0;
                e[g] = valid ? ptr[row_start + r + 1] : 0;
                acc[g] = 0.0;
            }
            bool more = true;
            while (more)
            {
                int c[RPG];
                double v[RPG];
#pragma unroll
                for (int g = 0; g < RPG; g++)
                {
                    const bool on = j[g] < e[g];
                    c[g] = on ? col[j[g]] : 0;
                    v[g] = on ? val[j[g]] : 0.0;
                }
                more = false;
#pragma unroll
                for (int g = 0; g < RPG; g++)
                {
                    if (j[g] < e[g]) acc[g] += v[g] * x[c[g]];
                    j[g] += TPR;
                    more |= j[g] < e[g];
                }
            }
#pragma unroll
            for (int g = 0; g < RPG; g++)
            {
#pragma unroll
                for (int o = TPR / 2; o > 0; o >>= 1) acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], o, TPR);
                const long long r = r0 + g * RPW + sub;
                if (r < num_rows && lane == 0) epi(row_start + (int)r, acc[g]);
            }
        }
    }
}

// flat variant for rows longer than a few entries: a warp owns R consecutive rows, i.e. one contiguous
// slice of col/val.  Lanes stride FLAT through the slice (fully coalesced, every lane busy whatever
// the individual row lengths, independent iterations -> many loads in flight), park the products
// val * x[col] in the warp's shared-memory strip, and 32/R lanes per row then add each row's
// products in a fixed order.  No CTA-wide barrier; a slice longer than the strip falls back to the
// sub-warp-per-row walk for that warp.
constexpr int kFlatCap = 512; // products per warp strip

template <int R, class Epi>
__global__ void __launch_bounds__(kSpThreads) k_spmv_flat(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ x, int row_start, int num_rows, Epi epi)
{
    constexpr int TPR = 32 / R;
    __shared__ double sprod[kSpThreads / 32][kFlatCap];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / TPR, sl = lane % TPR;
    double *strip = sprod[warp];
    const long long rows_per_grid = (long long)gridDim.x * (kSpThreads / 32) * R;
    for (long long r0 = ((long long)blockIdx.x * (kSpThreads / 32) + warp) * R; r0 < num_rows; r0 += rows_per_grid)
    {
        const int pl0 = ptr[row_start + min(r0 + lane, (long long)num_rows)];
        const int pl1 = ptr[row_start + min(r0 + lane + 1, (long long)num_rows)];
        const int p_begin = __shfl_sync(0xffffffffu, pl0, 0);
        const int n = __shfl_sync(0xffffffffu, pl1, R - 1) - p_begin;
        const int s = __shfl_sync(0xffffffffu, pl0, sub) - p_begin, e = __shfl_sync(0xffffffffu, pl1, sub) - p_begin;
        double acc = 0.0;
        if (n <= kFlatCap)
        {
            const int *c = col + p_begin;
            const double *v = val + p_begin;
            // batches of 8 x 32 entries: all col/val loads of a batch are issued before the first gather, all gathers
            // before the first store, so one batch costs two memory latencies whatever its size
            for (int base = 0; base < n; base += 256)
            {
                int cc[8];
                double vv[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                {
                    const int k = base + lane + 32 * u;
                    const bool on = k < n;
                    cc[u] = on ? c[k] : -1;
                    vv[u] = on ? v[k] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; u++) vv[u] *= cc[u] >= 0 ? x[cc[u]] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; u++)
                {
                    const int k = base + lane + 32 * u;
                    if (k < n) strip[k] = vv[u];
                }
            }
            __syncwarp();
            for (int j = s + sl; j < e; j += TPR) acc += strip[j];
            __syncwarp();
        }
        else
        {
            for (int j = p_begin + s + sl; j < p_begin + e; j += TPR) acc += val[j] * x[col[j]];
        }
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, TPR);
        if (sl == 0 && r0 + sub < num_rows) epi(row_start + (int)(r0 + sub), acc);
    }
}

template <int R, class Epi>
static void launch_flat(const int *ptr, const int *col, const double *val, const double *x, int row_start, int num_rows, cudaStream_t st, Epi epi)
{
    const int grid = stream_grid(((long long)num_rows + R - 1) / R * 32, kSpThreads, 1, 16);
    k_spmv_flat<R><<<grid, kSpThreads, 0, st>>>(ptr, col, val, x, row_start, num_rows, epi);
}

static int spmv_rpg()
{
    static const int v = [] { const char *e = getenv("PRFDD_SPMV_RPG"); const int r = e ? atoi(e) : 1; return (r == 2 || r == 4) ? r : 1; }();
    return v;
}

template <int TPR, class Epi>
static void launch_spmv(const int *ptr, const int *col, const double *val, const double *x, int row_start, int num_rows, cudaStream_t st, Epi epi)
{
    const int rpg = TPR == 1 ? 1 : spmv_rpg();
    const int grid = stream_grid((long long)num_rows * TPR / rpg, kSpThreads, 1, 16);
    if (rpg == 2) k_spmv<TPR, 2><<<grid, kSpThreads, 0, st>>>(ptr, col, val, x, row_start, num_rows, epi);
    else if (rpg == 4) k_spmv<TPR, 4><<<grid, kSpThreads, 0, st>>>(ptr, col, val, x, row_start, num_rows, epi);
    else k_spmv<TPR, 1><<<grid, kSpThreads, 0, st>>>(ptr, col, val, x, row_start, num_rows, epi);
}

template <class Epi>
static int spmv(const int *ptr, const int *col, const double *val, const double *x, int row_start, int num_rows, int tpr, cudaStream_t st, Epi epi)
{
    if (num_rows <= 0) return 0;
    if (tpr == 0) tpr = 4;
    if (tpr < 0 && !x) tpr = 1; // no product to form (u = 0 shortcut): the epilogue alone
    switch (tpr)
    {
    case -1: launch_flat<1>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case -2: launch_flat<2>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case -4: launch_flat<4>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case -8: launch_flat<8>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case -16: launch_flat<16>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case -32: launch_flat<32>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case 1: launch_spmv<1>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case 2: launch_spmv<2>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case 4: launch_spmv<4>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case 8: launch_spmv<8>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case 16: launch_spmv<16>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    case 32: launch_spmv<32>(ptr, col, val, x, row_start, num_rows, st, epi); break;
    default: return -6;
    }
    return launched();
}

// ---------------------------------------------------------------------------------------------
// streamed SpMV for the large AMG levels: one CTA owns a block of consecutive rows whose entries
// (<= CH of them) are contiguous in col/val.  One thread moves the block's col and val slices into
// shared memory with two cp.async.bulk copies completing on an mbarrier (the streaming part of the
// SpMV is then a deep DMA instead of per-lane loads that wait on each other); every thread then
// multiplies entries by the gathered x[col] flat over the block (all lanes busy whatever the row
// lengths, many independent gathers in flight per thread), and sub-warps of TPR lanes add the
// products of each row from shared memory in a fixed order and run the epilogue.
// row_blocks[b] .. row_blocks[b+1] are the rows of block b (prfdd_csr_row_blocks builds them).
// ---------------------------------------------------------------------------------------------
constexpr int kStThreads = 256;
constexpr int kStMaxRows = 1024; // rows per block (shared row-pointer slice)

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int CH, int TPR, class Epi>
__global__ void __launch_bounds__(kStThreads) k_spmv_stream(const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ val, const double *__restrict__ x, const int *__restrict__ row_blocks, int nnz_total, Epi epi)
{
    extern __shared__ __align__(16) unsigned char stream_smem[];
    double *sval = reinterpret_cast<double *>(stream_smem);
    int *scol = reinterpret_cast<int *>(stream_smem + sizeof(double) * CH);
    int *sptr = scol + CH;
    __shared__ __align__(8) unsigned long long bar;

    const int tid = threadIdx.x;
    const int r0 = row_blocks[blockIdx.x], r1 = row_blocks[blockIdx.x + 1], nr = r1 - r0;
    const int p0 = ptr[r0], p1 = ptr[r1];
    const int a0 = p0 & ~3;                                  // 16-byte aligned start of the col slice
    const int nb = max(min((p1 + 3) & ~3, nnz_total & ~3) - a0, 0); // entries moved by the bulk copies
    if (tid == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (nb > 0)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(12u * (uint32_t)nb) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(scol)), "l"(col + a0), "r"(4u * (uint32_t)nb), "r"(smem_addr(&bar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sval)), "l"(val + a0), "r"(8u * (uint32_t)nb), "r"(smem_addr(&bar)) : "memory");
        }
    }
    for (int i = tid; i <= nr; i += kStThreads) sptr[i] = ptr[r0 + i] - a0;
    for (int i = a0 + nb + tid; i < p1; i += kStThreads) // the < 4 entries at the very end of the matrix that no aligned copy covers
    {
        scol[i - a0] = col[i];
        sval[i - a0] = val[i];
    }
    __syncthreads();
    if (nb > 0)
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(&bar)), "r"(0u) : "memory");
    }
    // flat gather-multiply over the block's entries
    const int lo = p0 - a0, hi = p1 - a0;
    int i = lo + tid;
    for (; i + 3 * kStThreads < hi; i += 4 * kStThreads)
    {
        const double x0 = x[scol[i]], x1 = x[scol[i + kStThreads]], x2 = x[scol[i + 2 * kStThreads]], x3 = x[scol[i + 3 * kStThreads]];
        sval[i] *= x0;
        sval[i + kStThreads] *= x1;
        sval[i + 2 * kStThreads] *= x2;
        sval[i + 3 * kStThreads] *= x3;
    }
    for (; i < hi; i += kStThreads) sval[i] *= x[scol[i]];
    __syncthreads();
    // per-row sums
    constexpr int RPP = kStThreads / TPR;
    const int lane = tid % TPR;
    for (int base = 0; base < nr; base += RPP)
    {
        const int rl = base + tid / TPR;
        const bool valid = rl < nr;
        const int s = valid ? sptr[rl] : 0, e = valid ? sptr[rl + 1] : 0;
        double acc = 0.0;
        for (int j = s + lane; j < e; j += TPR) acc += sval[j];
#pragma unroll
        for (int o = TPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, TPR);
        if (valid && lane == 0) epi(r0 + rl, acc);
    }
}

constexpr size_t stream_smem_bytes(int ch) { return (size_t)ch * 12 + (kStMaxRows + 1) * 4; }
static int stream_chunk()
{
    static const int ch = [] { const char *e = getenv("PRFDD_STREAM_CHUNK"); const int v = e ? atoi(e) : 4096; return (v == 1024 || v == 2048) ? v : 4096; }();
    return ch;
}

template <int CH, int TPR, class Epi>
static void launch_stream_ch(const int *ptr, const int *col, const double *val, const double *x, const int *row_blocks, int num_blocks, int nnz_total, cudaStream_t st, Epi epi)
{
    auto kern = k_spmv_stream<CH, TPR, Epi>;
    static bool configured = false; // one per template instance
    if (!configured)
    {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem_bytes(CH));
        configured = true;
    }
    kern<<<num_blocks, kStThreads, stream_smem_bytes(CH), st>>>(ptr, col, val, x, row_blocks, nnz_total, epi);
}

template <int TPR, class Epi>
static void launch_stream(const int *ptr, const int *col, const double *val, const double *x, const int *row_blocks, int num_blocks, int nnz_total, cudaStream_t st, Epi epi)
{
    switch (stream_chunk())
    {
    case 1024: launch_stream_ch<1024, TPR>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    case 2048: launch_stream_ch<2048, TPR>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    default: launch_stream_ch<4096, TPR>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    }
}

template <class Epi>
static int spmv_stream(const int *ptr, const int *col, const double *val, const double *x, const int *row_blocks, int num_blocks, int nnz_total, int tpr, cudaStream_t st, Epi epi)
{
    if (num_blocks <= 0) return 0;
    if (((uintptr_t)col & 15) || ((uintptr_t)val & 15)) return -8; // bulk copies need 16-byte aligned arrays
    switch (tpr)
    {
    case 1: launch_stream<1>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    case 2: launch_stream<2>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    case 4: launch_stream<4>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    case 8: launch_stream<8>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    case 16: launch_stream<16>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    case 32: launch_stream<32>(ptr, col, val, x, row_blocks, num_blocks, nnz_total, st, epi); break;
    default: return -6;
    }
    return launched();
}

// ---------------------------------------------------------------------------------------------
// gather / scatter (index-map Q^T / Q) and halo
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather(double *__restrict__ nodes, const int *__restrict__ ptr, const int *__restrict__ col, const double *__restrict__ u, const double *__restrict__ weight, int num_nodes)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < num_nodes; v += stride)
    {
        const int s = ptr[v], e = ptr[v + 1];
        double acc = 0.0;
        for (int j = s; j < e; j++) acc += u[col[j]];
        nodes[v] = weight ? acc * weight[v] : acc;
    }
}

__global__ void __launch_bounds__(256) k_scatter(double *__restrict__ out, const int *__restrict__ node_of_point, const double *__restrict__ nodes, const double *__restrict__ mask, long long num_points)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < num_points; p += stride)
    {
        const double v = nodes[node_of_point[p]];
        out[p] = mask ? v * mask[p] : v;
    }
}

__global__ void __launch_bounds__(256) k_pack(double *__restrict__ buf, const double *__restrict__ nodes, const int *__restrict__ idx, int count)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) buf[i] = nodes[idx[i]];
}

__global__ void __launch_bounds__(256) k_unpack_add(double *__restrict__ nodes, const double *__restrict__ buf, const int *__restrict__ idx, int count)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) nodes[idx[i]] += buf[i];
}

__global__ void __launch_bounds__(256) k_scatter_assign(double *__restrict__ dst, const double *__restrict__ buf, const int *__restrict__ idx, int count)
{
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[idx[i]] = buf[i];
}

__global__ void __launch_bounds__(256) k_dense_solve(double *__restrict__ x, const double *__restrict__ Ainv, const double *__restrict__ b, int n)
{
    extern __shared__ double sb[];
    for (int t = threadIdx.x; t < n; t += blockDim.x) sb[t] = b[t];
    __syncthreads();
    for (int r = threadIdx.x; r < n; r += blockDim.x)
    {
        double acc = 0.0;
        for (int c = 0; c < n; c++) acc += Ainv[(size_t)r * n + c] * sb[c];
        x[r] = acc;
    }
}
} // namespace prfdd

using namespace prfdd;

extern "C" {

int prfdd_gather(double *nodes, const int *ptr, const int *col, const double *u, const double *weight, int num_nodes, prfdd_stream_t stream)
{
    if (num_nodes <= 0) return 0;
    k_gather<<<stream_grid(num_nodes, 256, 1, 16), 256, 0, S(stream)>>>(nodes, ptr, col, u, weight, num_nodes);
    return launched();
}

int prfdd_scatter(double *out, const int *node_of_point, const double *nodes, const double *mask, int num_points, prfdd_stream_t stream)
{
    if (num_points <= 0) return 0;
    k_scatter<<<stream_grid(num_points, 256, 2, 8), 256, 0, S(stream)>>>(out, node_of_point, nodes, mask, num_points);
    return launched();
}

int prfdd_halo_pack(double *buf, const double *nodes, const int *idx, int count, prfdd_stream_t stream)
{
    if (count <= 0) return 0;
    k_pack<<<stream_grid(count, 256, 1, 8), 256, 0, S(stream)>>>(buf, nodes, idx, count);
    return launched();
}

int prfdd_halo_unpack_add(double *nodes, const double *buf, const int *idx, int count, prfdd_stream_t stream)
{
    if (count <= 0) return 0;
    k_unpack_add<<<stream_grid(count, 256, 1, 8), 256, 0, S(stream)>>>(nodes, buf, idx, count);
    return launched();
}

int prfdd_scatter_assign(double *dst, const double *buf, const int *idx, int count, prfdd_stream_t stream)
{
    if (count <= 0) return 0;
    k_scatter_assign<<<stream_grid(count, 256, 1, 8), 256, 0, S(stream)>>>(dst, buf, idx, count);
    return launched();
}

int prfdd_csr_multiply(double *Au, const int *ptr, const int *col, const double *val, const double *u, int num_rows, int tpr, prfdd_stream_t stream)
{
    return spmv(ptr, col, val, u, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) { Au[row] = ax; });
}

int prfdd_csr_multiply_range(double *Au, const int *ptr, const int *col, const double *val, const double *u, int row_start, int row_end, int tpr, prfdd_stream_t stream)
{
    if (row_end < row_start) return -7; // csr_matrix.tpp:319-323
    return spmv(ptr, col, val, u, row_start, row_end - row_start + 1, tpr, S(stream), [=] __device__(int row, double ax) { Au[row] = ax; });
}

int prfdd_csr_multiply_weight(double *Au, const int *ptr, const int *col, const double *val, const double *u, const double *weight, int num_rows, int tpr, prfdd_stream_t stream)
{
    return spmv(ptr, col, val, u, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) { Au[row] = ax * weight[row]; });
}

int prfdd_csr_matvec(double *y, const int *ptr, const int *col, const double *val, const double *x, double alpha, double beta, int num_rows, int tpr, prfdd_stream_t stream)
{
    if (beta == 0.0)
        return spmv(ptr, col, val, x, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) { y[row] = alpha * ax; });
    return spmv(ptr, col, val, x, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) { y[row] = alpha * ax + beta * y[row]; });
}

int prfdd_csr_residual(double *v, const int *ptr, const int *col, const double *val, const double *u, const double *f, int num_rows, int tpr, prfdd_stream_t stream)
{
    return spmv(ptr, col, val, u, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) { v[row] = f[row] - ax; });
}

int prfdd_cheby_residual(double *r, double *t, const int *ptr, const int *col, const double *val, const double *u, const double *f, const double *ds, double c_hi, int num_rows, int tpr, prfdd_stream_t stream)
{
    // u == NULL: u = 0, the product A u is skipped (x == nullptr in k_spmv) and TPR is irrelevant
    return spmv(ptr, col, val, u, 0, num_rows, u ? tpr : 1, S(stream), [=] __device__(int row, double ax) {
        const double d = ds[row];
        const double rr = d * (f[row] - ax);
        r[row] = rr;
        t[row] = d * (c_hi * rr);
    });
}

int prfdd_cheby_step(double *u, double *t_out, const int *ptr, const int *col, const double *val, const double *t_in, const double *r, const double *ds, double c, int last, int u_is_zero, int num_rows, int tpr, prfdd_stream_t stream)
{
    if (last)
    {
        if (u_is_zero)
            return spmv(ptr, col, val, t_in, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) {
                const double d = ds[row];
                u[row] = d * (c * r[row] + d * ax);
            });
        return spmv(ptr, col, val, t_in, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) {
            const double d = ds[row];
            u[row] += d * (c * r[row] + d * ax);
        });
    }
    return spmv(ptr, col, val, t_in, 0, num_rows, tpr, S(stream), [=] __device__(int row, double ax) {
        const double d = ds[row];
        t_out[row] = d * (c * r[row] + d * ax);
    });
}

// streamed variants (k_spmv_stream) of the V-cycle's matrix passes; same arithmetic per row, the products of a row are
// added TPR-strided from shared memory
int prfdd_csr_row_blocks(const int *ptr_host, int num_rows, int *row_blocks, int *num_blocks)
{
    // greedy: consecutive rows while the block's entries (plus alignment slack) fit the shared-memory chunk
    int nb = 0, r = 0;
    while (r < num_rows)
    {
        int e = r;
        while (e < num_rows && e - r < kStMaxRows && ptr_host[e + 1] - ptr_host[r] <= stream_chunk() - 8) e++;
        if (e == r) return -9; // a single row does not fit: use the plain kernels for this matrix
        if (row_blocks) row_blocks[nb] = r;
        nb++;
        r = e;
    }
    if (row_blocks) row_blocks[nb] = num_rows;
    *num_blocks = nb;
    return 0;
}

int prfdd_csr_multiply_stream(double *Au, const int *ptr, const int *col, const double *val, const double *u, const int *row_blocks, int num_blocks, int nnz, int tpr, prfdd_stream_t stream)
{
    return spmv_stream(ptr, col, val, u, row_blocks, num_blocks, nnz, tpr, S(stream), [=] __device__(int row, double ax) { Au[row] = ax; });
}

int prfdd_csr_residual_stream(double *v, const int *ptr, const int *col, const double *val, const double *u, const double *f, const int *row_blocks, int num_blocks, int nnz, int tpr, prfdd_stream_t stream)
{
    return spmv_stream(ptr, col, val, u, row_blocks, num_blocks, nnz, tpr, S(stream), [=] __device__(int row, double ax) { v[row] = f[row] - ax; });
}

int prfdd_cheby_residual_stream(double *r, double *t, const int *ptr, const int *col, const double *val, const double *u, const double *f, const double *ds, double c_hi, const int *row_blocks, int num_blocks, int nnz, int tpr, prfdd_stream_t stream)
{
    return spmv_stream(ptr, col, val, u, row_blocks, num_blocks, nnz, tpr, S(stream), [=] __device__(int row, double ax) {
        const double d = ds[row];
        const double rr = d * (f[row] - ax);
        r[row] = rr;
        t[row] = d * (c_hi * rr);
    });
}

int prfdd_cheby_step_stream(double *u, double *t_out, const int *ptr, const int *col, const double *val, const double *t_in, const double *r, const double *ds, double c, int last, int u_is_zero, const int *row_blocks, int num_blocks, int nnz, int tpr, prfdd_stream_t stream)
{
    if (last)
    {
        if (u_is_zero)
            return spmv_stream(ptr, col, val, t_in, row_blocks, num_blocks, nnz, tpr, S(stream), [=] __device__(int row, double ax) {
                const double d = ds[row];
                u[row] = d * (c * r[row] + d * ax);
            });
        return spmv_stream(ptr, col, val, t_in, row_blocks, num_blocks, nnz, tpr, S(stream), [=] __device__(int row, double ax) {
            const double d = ds[row];
            u[row] += d * (c * r[row] + d * ax);
        });
    }
    return spmv_stream(ptr, col, val, t_in, row_blocks, num_blocks, nnz, tpr, S(stream), [=] __device__(int row, double ax) {
        const double d = ds[row];
        t_out[row] = d * (c * r[row] + d * ax);
    });
}

int prfdd_dense_solve(double *x, const double *Ainv, const double *b, int n, prfdd_stream_t stream)
{
    if (n <= 0) return 0;
    k_dense_solve<<<1, 256, sizeof(double) * n, S(stream)>>>(x, Ainv, b, n);
    return launched();
}

} // extern "C"
