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
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace prfdd
{
constexpr int kSpThreads = 256;

// generic SpMV with an epilogue functor: epi(row, Ax) executed by lane 0 of the row's sub-warp.
// The row loop is warp-uniform (all 32 lanes take the same number of trips) so the full-mask
// shuffles are always executed by the whole warp.  RPG > 1: every sub-warp walks RPG rows at once
// (rows r, r + 32/TPR, ...), which multiplies the independent col/val -> x load chains a lane has in
// flight; the long rows of the coarse AMG levels are latency bound without it.
// VT: value type of the matrix, the gathered vector and the row sums (double; float for the FP32 V-cycle, AMG/config.hpp:4)
template <int TPR, int RPG, bool UNIT, class VT, class Epi>
__global__ void __launch_bounds__(kSpThreads) k_spmv(const int *__restrict__ ptr, const int *__restrict__ col, const VT *__restrict__ val, const VT *__restrict__ x, int row_start, int num_rows, int skip_len, Epi epi)
{
    pdl_wait();
    constexpr int RPW = 32 / TPR; // rows per warp and pass
    const int lane = threadIdx.x % TPR;
    const int sub = (threadIdx.x & 31) / TPR;
    const long long r_begin = ((long long)blockIdx.x * (kSpThreads / 32) + (threadIdx.x >> 5)) * RPW * RPG;
    const long long r_end = num_rows;
    const long long r_step = (long long)gridDim.x * (kSpThreads / TPR) * RPG;
    for (long long r0 = r_begin; r0 < r_end; r0 += r_step)
    {
        int j[RPG], e[RPG];
        bool skip[RPG];
        VT acc[RPG];
#pragma unroll
        for (int g = 0; g < RPG; g++)
        {
            const long long r = r0 + g * RPW + sub;
            const bool valid = x && r < r_end;
            const int s = valid ? ptr[row_start + r] : 0;
            e[g] = valid ? ptr[row_start + r + 1] : 0;
            skip[g] = e[g] - s > skip_len; // a listed long row: k_spmv_long owns it
            if (skip[g]) e[g] = s;
            j[g] = s + lane;
            acc[g] = VT(0);
        }
        if constexpr (RPG == 1)
        {
            for (int k = j[0]; k < e[0]; k += TPR) acc[0] += UNIT ? x[col[k]] : val[k] * x[col[k]];
        }
        else
        {
            bool more = true;
            while (more)
            {
                int c[RPG];
                VT v[RPG];
#pragma unroll
                for (int g = 0; g < RPG; g++)
                {
                    const bool on = j[g] < e[g];
                    c[g] = on ? col[j[g]] : 0;
                    v[g] = (on && !UNIT) ? val[j[g]] : VT(1);
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
        }
#pragma unroll
        for (int g = 0; g < RPG; g++)
        {
#pragma unroll
            for (int o = TPR / 2; o > 0; o >>= 1) acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], o, TPR);
            const long long r = r0 + g * RPW + sub;
            if (r < r_end && lane == 0 && !skip[g]) epi(row_start + (int)r, acc[g]);
        }
    }
}

// Long rows.  A warp finishes when its longest row does, and under load one dependent col -> x step costs microseconds, so a
// thread that walks a 60- or 300-entry row alone (hanging-node rows of the non-conforming composite grid in the low-order
// matrix, and the columns of its Q that become rows of Q^T) outlasts the rest of the kernel.  A matrix descriptor can carry the
// list of its rows longer than a threshold (prfdd_csr_matrix::long_rows): the row-group kernels then skip those rows and
// k_spmv_long gives each a whole warp.  Same epilogue, fixed summation order.
template <bool UNIT, class VT, class Epi>
__global__ void __launch_bounds__(kSpThreads) k_spmv_long(const int *__restrict__ ptr, const int *__restrict__ col, const VT *__restrict__ val, const VT *__restrict__ x, const int *__restrict__ rows, int count, int num_rows, Epi epi)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (kSpThreads / 32) + (threadIdx.x >> 5);
    if (w >= count) return; // whole warps leave together
    const int row = rows[w];
    if (row >= num_rows) return;
    VT acc = VT(0);
    for (int j = ptr[row] + lane; j < ptr[row + 1]; j += 32) acc += UNIT ? x[col[j]] : val[j] * x[col[j]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) epi(row, acc);
}

// one entry per row (ptr == NULL): out = epi(row, [val] x[col[row]])  -- Q of a conforming region
template <bool UNIT, class VT, class Epi>
__global__ void __launch_bounds__(256) k_spmv_single(const int *__restrict__ col, const VT *__restrict__ val, const VT *__restrict__ x, int row_start, int num_rows, Epi epi)
{
    pdl_wait();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long r = row_start + (long long)blockIdx.x * blockDim.x + threadIdx.x; r < row_start + num_rows; r += stride) epi((int)r, UNIT ? x[col[r]] : val[r] * x[col[r]]);
}

// the descriptor as the kernels see it, for either value type
template <class VT>
struct CsrView
{
    const int *ptr, *col;
    const VT *val;
    int num_rows, num_cols, num_nnz, threads_per_row;
    const int *long_rows;
    int num_long_rows, long_row_threshold;
    const int *sell_off, *sell_col;
    const VT *sell_val;
    const int *sell_row;
    int sell_num_slices, sell_lanes, sell_window;
};
static CsrView<double> view(const prfdd_csr_matrix &A) { return {A.ptr, A.col, A.val, A.num_rows, A.num_cols, A.num_nnz, A.threads_per_row, A.long_rows, A.num_long_rows, A.long_row_threshold, A.sell_off, A.sell_col, A.sell_val, A.sell_row, A.sell_num_slices, A.sell_lanes, A.sell_window}; }
static CsrView<float> view(const prfdd_csr_matrix_f32 &A) { return {A.ptr, A.col, A.val, A.num_rows, A.num_cols, A.num_nnz, A.threads_per_row, A.long_rows, A.num_long_rows, A.long_row_threshold, A.sell_off, A.sell_col, A.sell_val, A.sell_row, A.sell_num_slices, A.sell_lanes, A.sell_window}; }

// Sliced layout (prfdd_sell_layout): a warp owns one slice of 32/T slots; the walk over the slice is warp-uniform, every col / val
// load of the warp is ONE contiguous line request, U chunks (col, val, then the gathers) are in flight per lane.  Lane t of a
// row takes its entries t, t + T, ... in order and the lanes are combined by the same shuffle tree as in k_spmv<T>: same sums.
// Scheduling note (measured, profiles/r2_notes.txt): ptxas fits this loop into 32 registers and issues its loads as dependent
// col/val -> x pairs; stating a larger budget in the launch bounds makes it issue U col / val loads and U gathers back to back
// (40-53 registers), which is SLOWER here (level 1: 46 -> 50-56 us) -- the kernel is bound by L1 wavefronts (one per distinct
// 128-byte line a warp load touches, ~2 cycles each), not by the latency of a warp's chain, and 64 resident warps hide that latency.
template <int T, int U, class VT, class Epi>
__global__ void __launch_bounds__(kSpThreads) k_spmv_sell(const int *__restrict__ off, const int *__restrict__ scol, const VT *__restrict__ sval, const int *__restrict__ srow, const VT *__restrict__ x, int num_slices, int num_rows, Epi epi)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (kSpThreads / 32);
    for (int s = blockIdx.x * (kSpThreads / 32) + (threadIdx.x >> 5); s < num_slices; s += nwarps)
    {
        // lanes 0 and 1 fetch the slice bounds in one request
        const int o = off[s + (lane & 1)];
        const int b = __shfl_sync(0xffffffffu, o, 0), e = __shfl_sync(0xffffffffu, o, 1);
        int row = -1;
        if (lane % T == 0)
        {
            const int slot = s * (32 / T) + lane / T;
            row = srow ? srow[slot] : (slot < num_rows ? slot : -1);
        }
        VT acc = VT(0);
        for (int k = b + lane; k < e; k += 32 * U)
        {
            int c[U];
            VT v[U];
#pragma unroll
            for (int i = 0; i < U; i++)
            {
                const bool on = k + 32 * i < e;
                c[i] = on ? scol[k + 32 * i] : 0;
                v[i] = on ? sval[k + 32 * i] : VT(0);
            }
#pragma unroll
            for (int i = 0; i < U; i++)
                if (k + 32 * i < e) acc += v[i] * x[c[i]];
        }
#pragma unroll
        for (int o2 = T / 2; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2, T);
        if (row >= 0) epi(row, acc);
    }
}

// Few slices (a wave or two of warps: the coarse AMG levels and their R / P): such a launch lasts as long as one warp's dependent
// chain, so here the loads of a trip ARE issued back to back -- plain pointer bumps, immediate offsets and a stated register
// budget make ptxas keep U col / val loads and then U gathers in flight (see the scheduling note above).  Same sums, same order.
template <int T, int U, class VT, class Epi>
__global__ void __launch_bounds__(kSpThreads, 4) k_spmv_sell_small(const int *__restrict__ off, const int *__restrict__ scol, const VT *__restrict__ sval, const int *__restrict__ srow, const VT *__restrict__ x, int num_slices, int num_rows, Epi epi)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (kSpThreads / 32);
    for (int s = blockIdx.x * (kSpThreads / 32) + (threadIdx.x >> 5); s < num_slices; s += nwarps)
    {
        const int o = off[s + (lane & 1)];
        const int b = __shfl_sync(0xffffffffu, o, 0), e = __shfl_sync(0xffffffffu, o, 1);
        int row = -1;
        if (lane % T == 0)
        {
            const int slot = s * (32 / T) + lane / T;
            row = srow ? srow[slot] : (slot < num_rows ? slot : -1);
        }
        VT acc = VT(0);
        const int *cp = scol + b + lane;
        const VT *vp = sval + b + lane;
        int n = (e - b) >> 5;
        while (n >= U)
        {
            int c[U];
            VT v[U], xv[U];
#pragma unroll
            for (int i = 0; i < U; i++) { c[i] = cp[32 * i]; v[i] = vp[32 * i]; }
#pragma unroll
            for (int i = 0; i < U; i++) xv[i] = x[c[i]];
#pragma unroll
            for (int i = 0; i < U; i++) acc += v[i] * xv[i];
            cp += 32 * U; vp += 32 * U; n -= U;
        }
        if (n > 0)
        {
            // the partial trip re-reads the lane's last chunk for the slots past the slice and adds it with value 0
            int c[U];
            VT v[U], xv[U];
#pragma unroll
            for (int i = 0; i < U - 1; i++) { const int q = i < n ? i : n - 1; c[i] = cp[32 * q]; v[i] = vp[32 * q]; }
#pragma unroll
            for (int i = 0; i < U - 1; i++) xv[i] = x[c[i]];
#pragma unroll
            for (int i = 0; i < U - 1; i++) acc += (i < n ? v[i] : VT(0)) * xv[i];
        }
#pragma unroll
        for (int o2 = T / 2; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2, T);
        if (row >= 0) epi(row, acc);
    }
}

// Sorted layout (sell_row set): the rows of a slice are scattered over their 256-row window, so row-indexed epilogue operands
// (ds, r, f, the outputs) would cost a line request per row.  One CTA owns one window: its warps walk the window's slices (dealt
// round robin: they are sorted by width), park the row sums in shared memory, and after one barrier thread t finishes row
// w0 + t -- every epilogue access is contiguous again.  sell_row holds the row of each slot; only its offset in the window is used.
template <int T, int U, class VT, class Epi>
__global__ void __launch_bounds__(kSpThreads) k_spmv_sell_window(const int *__restrict__ off, const int *__restrict__ scol, const VT *__restrict__ sval, const int *__restrict__ srow, const VT *__restrict__ x, int num_slices, int num_rows, Epi epi)
{
    pdl_wait();
    constexpr int R = 32 / T;          // rows per slice
    constexpr int SPW = kSpThreads / R; // slices per window of kSpThreads rows
    __shared__ VT ax[kSpThreads];
    const int lane = threadIdx.x & 31;
    const int w0 = blockIdx.x * kSpThreads;
    for (int q = threadIdx.x >> 5; q < SPW; q += kSpThreads / 32)
    {
        const int s = blockIdx.x * SPW + q;
        if (s >= num_slices) break; // warp-uniform
        const int o = off[s + (lane & 1)];
        const int b = __shfl_sync(0xffffffffu, o, 0), e = __shfl_sync(0xffffffffu, o, 1);
        int row = -1;
        if (lane % T == 0) row = srow[s * R + lane / T];
        VT acc = VT(0);
        for (int k = b + lane; k < e; k += 32 * U)
        {
            int c[U];
            VT v[U];
#pragma unroll
            for (int i = 0; i < U; i++)
            {
                const bool on = k + 32 * i < e;
                c[i] = on ? scol[k + 32 * i] : 0;
                v[i] = on ? sval[k + 32 * i] : VT(0);
            }
#pragma unroll
            for (int i = 0; i < U; i++)
                if (k + 32 * i < e) acc += v[i] * x[c[i]];
        }
#pragma unroll
        for (int o2 = T / 2; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2, T);
        if (row >= 0) ax[row - w0] = acc;
    }
    __syncthreads();
    const int row = w0 + threadIdx.x;
    if (row < num_rows) epi(row, ax[threadIdx.x]);
}

template <int T, class VT, class Epi>
static void launch_sell(const CsrView<VT> &A, const VT *x, cudaStream_t st, Epi epi)
{
    static const int waves = getenv("PRFDD_SELL_WAVES") ? atoi(getenv("PRFDD_SELL_WAVES")) : 32; // CTAs per SM of the grid (measured on c2: 4 -> 15.0 ms, 8 -> 12.63, 16 -> 12.54, 32 -> 12.44, unlimited -> 12.50)
    const int grid = stream_grid(A.sell_num_slices, kSpThreads / 32, 1, waves);
    static const bool no_window = getenv("PRFDD_SELL_NO_WINDOW_KERNEL") != nullptr;
    static const bool no_small = getenv("PRFDD_SELL_NO_SMALL_KERNEL") != nullptr;
    static const int small_max = getenv("PRFDD_SELL_SMALL_MAX") ? atoi(getenv("PRFDD_SELL_SMALL_MAX")) : 4 * (kSpThreads / 32) * num_sms();
    // sorted inside windows of kSpThreads rows (prfdd_sell_layout with window_rows = 256): CTA per window when that still fills the chip
    // twice over and a warp has at most two slices to walk (measured on the c2 hierarchy: level-1 A 47.2 -> 44.5 us, R of level 0
    // 27.4 -> 21.0 us; but 135 k rows x 8 lanes 29.6 -> 35.8 us, 18 k rows x 16 lanes 12.8 -> 43 us)
    if (A.sell_row && A.sell_window == kSpThreads && T <= 2 && A.num_rows >= 2 * 8 * kSpThreads * num_sms() && !no_window)
        launch_pdl(k_spmv_sell_window<T, 4, VT, Epi>, (A.num_rows + kSpThreads - 1) / kSpThreads, kSpThreads, 0, st, A.sell_off, A.sell_col, A.sell_val, A.sell_row, x, A.sell_num_slices, A.num_rows, epi);
    else if (A.sell_num_slices <= small_max && !no_small) // half a wave of warps (measured: at a full wave the plain kernel is as fast or faster)
        launch_pdl(k_spmv_sell_small<T, 4, VT, Epi>, grid, kSpThreads, 0, st, A.sell_off, A.sell_col, A.sell_val, A.sell_row, x, A.sell_num_slices, A.num_rows, epi);
    else
        launch_pdl(k_spmv_sell<T, 4, VT, Epi>, grid, kSpThreads, 0, st, A.sell_off, A.sell_col, A.sell_val, A.sell_row, x, A.sell_num_slices, A.num_rows, epi);
}

template <int TPR, bool UNIT, class VT, class Epi>
static void launch_spmv(const CsrView<VT> &A, const VT *x, int row_start, int num_rows, bool lr, cudaStream_t st, Epi epi)
{
    // two rows per sub-warp once the matrix is large enough to fill the machine with half as many warps
    // (AMG level 1 of the 16^3 N=7 problem, 27 entries/row: 54.9 -> 49.1 us; profiles/r1_notes.txt)
    const bool two = TPR >= 4 && (long long)num_rows * TPR >= (1ll << 21);
    const int skip_len = lr ? A.long_row_threshold : 0x7fffffff;
    static const int waves = getenv("PRFDD_SPMV_WAVES") ? atoi(getenv("PRFDD_SPMV_WAVES")) : 16; // experiment knob: CTAs per SM of the grid
    const int grid = stream_grid((long long)num_rows * TPR / (two ? 2 : 1), kSpThreads, 1, waves);
    if (two) launch_pdl(k_spmv<TPR, 2, UNIT, VT, Epi>, grid, kSpThreads, 0, st, A.ptr, A.col, A.val, x, row_start, num_rows, skip_len, epi);
    else launch_pdl(k_spmv<TPR, 1, UNIT, VT, Epi>, grid, kSpThreads, 0, st, A.ptr, A.col, A.val, x, row_start, num_rows, skip_len, epi);
    if (lr) launch_pdl(k_spmv_long<UNIT, VT, Epi>, (A.num_long_rows + kSpThreads / 32 - 1) / (kSpThreads / 32), kSpThreads, 0, st, A.ptr, A.col, A.val, x, A.long_rows, A.num_long_rows, num_rows, epi);
}

// dispatch on the descriptor.  epi_bytes: algorithmic bytes per row of the epilogue's operands
template <class VT, class Epi>
static int spmv(const CsrView<VT> &A, const VT *x, int row_start, int num_rows, cudaStream_t st, double epi_bytes, Epi epi)
{
    if (num_rows <= 0) return 0;
    if (!A.col) return -8;
    const bool unit = A.val == nullptr;
    // algorithmic bytes: col 4 (+ val) per entry, ptr 4 per row, the gathered vector once, the epilogue's operands
    double bytes = epi_bytes * num_rows;
    if (x)
    {
        const double entries = !A.ptr ? (double)num_rows : (num_rows == A.num_rows && A.num_nnz >= 0) ? (double)A.num_nnz : 0.0;
        bytes += (unit ? 4.0 : 4.0 + sizeof(VT)) * entries + (A.ptr ? 4.0 * (num_rows + 1) : 0.0) + (double)sizeof(VT) * (A.num_cols > 0 ? A.num_cols : num_rows);
    }
    if (!A.ptr)
    {
        // one entry per row
        if (!x) return -8;
        if (unit) launch_pdl(k_spmv_single<true, VT, Epi>, stream_grid(num_rows, 256, 2, 8), 256, 0, st, A.col, nullptr, x, row_start, num_rows, epi);
        else launch_pdl(k_spmv_single<false, VT, Epi>, stream_grid(num_rows, 256, 2, 8), 256, 0, st, A.col, A.val, x, row_start, num_rows, epi);
        return launched(bytes);
    }
    if (A.sell_col && x && !unit && row_start == 0 && num_rows == A.num_rows)
    {
        switch (A.sell_lanes)
        {
        case 1: launch_sell<1, VT>(A, x, st, epi); break;
        case 2: launch_sell<2, VT>(A, x, st, epi); break;
        case 4: launch_sell<4, VT>(A, x, st, epi); break;
        case 8: launch_sell<8, VT>(A, x, st, epi); break;
        case 16: launch_sell<16, VT>(A, x, st, epi); break;
        case 32: launch_sell<32, VT>(A, x, st, epi); break;
        default: return -6;
        }
        if (A.num_long_rows > 0 && A.long_rows)
        {
            // the listed rows are not in the sliced copy (slot row -1): warp per row from the CSR arrays, as after the row kernels
            launch_pdl(k_spmv_long<false, VT, Epi>, (A.num_long_rows + kSpThreads / 32 - 1) / (kSpThreads / 32), kSpThreads, 0, st, A.ptr, A.col, A.val, x, A.long_rows, A.num_long_rows, num_rows, epi);
            prfdd_launch_count_add(1);
        }
        return launched(bytes);
    }
    const int tpr = A.threads_per_row > 0 ? A.threads_per_row : 4;
    const bool lr = x && row_start == 0 && tpr < 32 && A.num_long_rows > 0 && A.long_rows;
    if (unit)
    {
        // matrices without values are Q^T-like: short rows
        switch (tpr)
        {
        case 1: launch_spmv<1, true, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 2: launch_spmv<2, true, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 4: launch_spmv<4, true, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 8: launch_spmv<8, true, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        default: return -6;
        }
    }
    else
    {
        switch (tpr)
        {
        case 1: launch_spmv<1, false, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 2: launch_spmv<2, false, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 4: launch_spmv<4, false, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 8: launch_spmv<8, false, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 16: launch_spmv<16, false, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        case 32: launch_spmv<32, false, VT>(A, x, row_start, num_rows, lr, st, epi); break;
        default: return -6;
        }
    }
    if (lr) prfdd_launch_count_add(1);
    return launched(bytes);
}

// ---- the SpMV family, once for both value types -----------------------------------------------------------------------
template <class VT>
static int t_multiply(VT *Au, const CsrView<VT> &A, const VT *u, cudaStream_t st)
{
    return spmv(A, u, 0, A.num_rows, st, 1.0 * sizeof(VT), [=] __device__(int row, VT ax) { Au[row] = ax; });
}
template <class VT>
static int t_multiply_range(VT *Au, const CsrView<VT> &A, const VT *u, int row_start, int row_end, cudaStream_t st)
{
    if (row_end < row_start) return -7; // csr_matrix.tpp:319-323
    return spmv(A, u, row_start, row_end - row_start + 1, st, 1.0 * sizeof(VT), [=] __device__(int row, VT ax) { Au[row] = ax; });
}
template <class VT>
static int t_multiply_weight(VT *Au, const CsrView<VT> &A, const VT *u, const VT *weight, cudaStream_t st)
{
    return spmv(A, u, 0, A.num_rows, st, 2.0 * sizeof(VT), [=] __device__(int row, VT ax) { Au[row] = ax * weight[row]; });
}
template <class VT>
static int t_matvec(VT *y, const CsrView<VT> &A, const VT *x, VT alpha, VT beta, cudaStream_t st)
{
    if (beta == VT(0))
        return spmv(A, x, 0, A.num_rows, st, 1.0 * sizeof(VT), [=] __device__(int row, VT ax) { y[row] = alpha * ax; });
    return spmv(A, x, 0, A.num_rows, st, 2.0 * sizeof(VT), [=] __device__(int row, VT ax) { y[row] = alpha * ax + beta * y[row]; });
}
template <class VT>
static int t_residual(VT *v, const CsrView<VT> &A, const VT *u, const VT *f, cudaStream_t st)
{
    return spmv(A, u, 0, A.num_rows, st, 2.0 * sizeof(VT), [=] __device__(int row, VT ax) { v[row] = f[row] - ax; });
}
template <class VT>
static int t_cheby_residual(VT *r, VT *t, const CsrView<VT> &A, const VT *u, const VT *f, const VT *ds, VT c_hi, cudaStream_t st)
{
    // u == NULL: u = 0, the product A u is skipped (x == nullptr in k_spmv) and the launch shape is irrelevant
    CsrView<VT> B = A;
    if (!u) B.threads_per_row = 1;
    return spmv(B, u, 0, A.num_rows, st, 4.0 * sizeof(VT), [=] __device__(int row, VT ax) {
        const VT d = ds[row];
        const VT rr = d * (f[row] - ax);
        r[row] = rr;
        t[row] = d * (c_hi * rr);
    });
}
template <class VT>
static int t_restrict_cheby_residual(VT *f, VT *r, VT *t, const CsrView<VT> &R, const VT *v, const VT *ds, VT c_hi, cudaStream_t st)
{
    // f = R v fused with the zero-guess head of the coarse level's smoothing: r = ds f, t = ds (c_hi r)
    return spmv(R, v, 0, R.num_rows, st, 4.0 * sizeof(VT), [=] __device__(int row, VT ax) {
        const VT d = ds[row];
        const VT rr = d * ax;
        f[row] = ax;
        r[row] = rr;
        t[row] = d * (c_hi * rr);
    });
}
template <class VT>
static int t_cheby_step(VT *u, VT *t_out, const CsrView<VT> &A, const VT *t_in, const VT *r, const VT *ds, VT c, int last, int u_is_zero, cudaStream_t st)
{
    if (last)
    {
        if (u_is_zero)
            return spmv(A, t_in, 0, A.num_rows, st, 3.0 * sizeof(VT), [=] __device__(int row, VT ax) {
                const VT d = ds[row];
                u[row] = d * (c * r[row] + d * ax);
            });
        return spmv(A, t_in, 0, A.num_rows, st, 4.0 * sizeof(VT), [=] __device__(int row, VT ax) {
            const VT d = ds[row];
            u[row] += d * (c * r[row] + d * ax);
        });
    }
    return spmv(A, t_in, 0, A.num_rows, st, 3.0 * sizeof(VT), [=] __device__(int row, VT ax) {
        const VT d = ds[row];
        t_out[row] = d * (c * r[row] + d * ax);
    });
}

// dense x = M b for both value types (k_dense_solve above is the double instance)
template <class VT>
__global__ void __launch_bounds__(256, 4) k_dense_t(VT *__restrict__ x, const VT *__restrict__ M, const VT *__restrict__ b, int n)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return; // whole warps leave together
    const VT *__restrict__ row = M + (size_t)r * n;
    // four independent partial sums per lane: the eight loads of a trip are in flight together (the matrix is L2 resident, the
    // launch is one wave: its length is one warp's chain)
    VT a0 = VT(0), a1 = VT(0), a2 = VT(0), a3 = VT(0);
    int c = lane;
    for (; c + 96 < n; c += 128)
    {
        const VT m0 = row[c], m1 = row[c + 32], m2 = row[c + 64], m3 = row[c + 96];
        const VT b0 = b[c], b1 = b[c + 32], b2 = b[c + 64], b3 = b[c + 96];
        a0 += m0 * b0; a1 += m1 * b1; a2 += m2 * b2; a3 += m3 * b3;
    }
    for (; c < n; c += 32) a0 += row[c] * b[c];
    VT acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) x[r] = acc;
}

template <class A, class B>
__global__ void __launch_bounds__(256) k_cast(A *__restrict__ dst, const B *__restrict__ src, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (A)src[i];
}

// the pointer-argument entry points (the reference's kernel signatures + one lanes-per-row hint): no long-row list, no staging
static prfdd_csr_matrix plain(const int *ptr, const int *col, const double *val, int num_rows, int tpr)
{
    prfdd_csr_matrix A = {};
    A.ptr = ptr; A.col = col; A.val = val;
    A.num_rows = num_rows; A.num_nnz = -1;
    A.threads_per_row = tpr <= 0 ? 4 : tpr;
    return A;
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

// dense x = M b, one warp per row, lanes stride through the row (coalesced), fixed-order shuffle sum.  Serves the coarsest-level
// inverse (a handful of rows) and the collapsed coarse levels of the V-cycle (a couple of thousand rows, amg.hpp).
__global__ void __launch_bounds__(256, 4) k_dense_solve(double *__restrict__ x, const double *__restrict__ Ainv, const double *__restrict__ b, int n)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return; // whole warps leave together
    const double *__restrict__ row = Ainv + (size_t)r * n;
    // four independent partial sums per lane (see k_dense_t)
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int c = lane;
    for (; c + 96 < n; c += 128)
    {
        const double m0 = row[c], m1 = row[c + 32], m2 = row[c + 64], m3 = row[c + 96];
        const double b0 = b[c], b1 = b[c + 32], b2 = b[c + 64], b3 = b[c + 96];
        a0 += m0 * b0; a1 += m1 * b1; a2 += m2 * b2; a3 += m3 * b3;
    }
    for (; c < n; c += 32) a0 += row[c] * b[c];
    double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) x[r] = acc;
}
} // namespace prfdd

using namespace prfdd;

template <class VT>
static int sell_fill(const int *ptr, const int *col, const double *val, int num_rows, int lanes, const int *off, const int *slot_row, int *scol, VT *sval)
{
    if (!ptr || !col || !val || !off || !slot_row || !scol || !sval) return -8;
    const int R = 32 / lanes;
    const int num_slices = (num_rows + R - 1) / R;
    for (int s = 0; s < num_slices; s++)
    {
        const int width = (off[s + 1] - off[s]) / 32;
        for (int q = 0; q < R; q++)
        {
            const int r = slot_row[s * R + q];
            const int b = r >= 0 ? ptr[r] : 0, len = r >= 0 ? ptr[r + 1] - ptr[r] : 0;
            const int pad_col = len > 0 ? col[b + len - 1] : 0; // padding re-reads the row's last column: no new line is touched
            for (int k = 0; k < width; k++)
                for (int t = 0; t < lanes; t++)
                {
                    const int i = k * lanes + t;
                    const size_t pos = (size_t)off[s] + 32 * (size_t)k + q * lanes + t;
                    scol[pos] = i < len ? col[b + i] : pad_col;
                    sval[pos] = i < len ? (VT)val[b + i] : VT(0);
                }
        }
    }
    return 0;
}

extern "C" {

int prfdd_gather(double *nodes, const int *ptr, const int *col, const double *u, const double *weight, int num_nodes, prfdd_stream_t stream)
{
    if (num_nodes <= 0) return 0;
    k_gather<<<stream_grid(num_nodes, 256, 1, 16), 256, 0, S(stream)>>>(nodes, ptr, col, u, weight, num_nodes);
    return launched((weight ? 20.0 : 12.0) * num_nodes); // ptr 4 + result 8 (+ weight 8) per node; the 12 bytes per gathered point are not known here (the caller adds them)
}

int prfdd_scatter(double *out, const int *node_of_point, const double *nodes, const double *mask, int num_points, prfdd_stream_t stream)
{
    if (num_points <= 0) return 0;
    k_scatter<<<stream_grid(num_points, 256, 2, 8), 256, 0, S(stream)>>>(out, node_of_point, nodes, mask, num_points);
    return launched((mask ? 20.0 : 12.0) * num_points); // index 4 + result 8 (+ mask 8) per point; the node vector (8 per node) is not known here
}

int prfdd_halo_pack(double *buf, const double *nodes, const int *idx, int count, prfdd_stream_t stream)
{
    if (count <= 0) return 0;
    k_pack<<<stream_grid(count, 256, 1, 8), 256, 0, S(stream)>>>(buf, nodes, idx, count);
    return launched(20.0 * count);
}

int prfdd_halo_unpack_add(double *nodes, const double *buf, const int *idx, int count, prfdd_stream_t stream)
{
    if (count <= 0) return 0;
    k_unpack_add<<<stream_grid(count, 256, 1, 8), 256, 0, S(stream)>>>(nodes, buf, idx, count);
    return launched(28.0 * count);
}

int prfdd_scatter_assign(double *dst, const double *buf, const int *idx, int count, prfdd_stream_t stream)
{
    if (count <= 0) return 0;
    k_scatter_assign<<<stream_grid(count, 256, 1, 8), 256, 0, S(stream)>>>(dst, buf, idx, count);
    return launched(20.0 * count);
}

// lanes per row from the average row length (measured on B200, profiles/r1_spmv_tpr.txt and r1_notes.txt)
// 17.8 entries/row (level 1 of the 2-rank composite hierarchy): 8 lanes x 2 rows beats 2 lanes (2-GPU solve 25.9 -> 25.1 ms), hence 14
static int lanes_per_row(double avg, int num_rows)
{
    static const double t1 = getenv("PRFDD_SPMV_T1") ? atof(getenv("PRFDD_SPMV_T1")) : 10.0; // experiment knobs
    static const double t2 = getenv("PRFDD_SPMV_T2") ? atof(getenv("PRFDD_SPMV_T2")) : 14.0;
    return avg <= t1 ? 1 : avg <= t2 ? 2 : (avg > 60 && num_rows < 50000) ? 16 : 8;
}

int prfdd_csr_plan(prfdd_csr_matrix *A, const int *ptr_host, int *long_rows_host, int capacity)
{
    if (!A || !ptr_host || A->num_rows < 0) return -8;
    static const bool no_long = getenv("PRFDD_SPMV_NO_LONG_ROWS") != nullptr;
    const int n = A->num_rows;
    const int nnz = ptr_host[n];
    A->num_nnz = nnz;
    A->long_rows = nullptr;
    A->num_long_rows = 0;
    A->long_row_threshold = 0;
    const double avg = (double)nnz / (n > 0 ? n : 1);
    const int tpr = lanes_per_row(avg, n);
    A->threads_per_row = tpr;
    if (n == 0 || nnz == 0) return 0;
    // rows much longer than the average (and than the lanes that walk them can absorb) go to the warp-per-row launch; worth a
    // second launch only when some row is several times the threshold (the tail it removes is then longer than the launch)
    const int threshold = std::max(std::max(16, 4 * tpr), (int)std::ceil(3.0 * avg));
    int count = 0, longest = 0;
    for (int r = 0; r < n; r++)
    {
        const int len = ptr_host[r + 1] - ptr_host[r];
        longest = std::max(longest, len);
        if (len > threshold) count++;
    }
    int listed = 0;
    if (!no_long && count > 0 && tpr < 32 && longest >= 4 * threshold && (double)count <= 0.05 * n)
    {
        if (count > capacity || !long_rows_host) return -count;
        for (int r = 0; r < n; r++)
            if (ptr_host[r + 1] - ptr_host[r] > threshold) long_rows_host[listed++] = r;
        A->long_row_threshold = threshold;
    }
    return listed;
}

long long prfdd_sell_layout(const int *ptr_host, int num_rows, int lanes, int window_rows, int *slice_off_host, int *slot_row_host)
{
    if (!ptr_host || !slice_off_host || !slot_row_host || num_rows < 0) return -8;
    if (lanes != 1 && lanes != 2 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32) return -6;
    const int R = 32 / lanes;
    const int num_slices = (num_rows + R - 1) / R;
    std::vector<int> order((size_t)num_rows);
    for (int r = 0; r < num_rows; r++) order[r] = r;
    auto chunks = [&](int r) { return (ptr_host[r + 1] - ptr_host[r] + lanes - 1) / lanes; };
    if (window_rows > R)
        for (int w0 = 0; w0 < num_rows; w0 += window_rows)
            std::stable_sort(order.begin() + w0, order.begin() + std::min(num_rows, w0 + window_rows), [&](int a, int b) { return chunks(a) > chunks(b); });
    long long total = 0;
    slice_off_host[0] = 0;
    for (int s = 0; s < num_slices; s++)
    {
        int width = 0;
        for (int q = 0; q < R; q++)
        {
            const int slot = s * R + q;
            const int r = slot < num_rows ? order[slot] : -1;
            slot_row_host[slot] = r;
            if (r >= 0) width = std::max(width, chunks(r));
        }
        total += 32ll * width;
        if (total > 0x7fffffffll) return -9;
        slice_off_host[s + 1] = (int)total;
    }
    return total;
}

int prfdd_sell_fill(const int *ptr_host, const int *col_host, const double *val_host, int num_rows, int lanes, const int *slice_off_host, const int *slot_row_host, int *sell_col_host, double *sell_val_host)
{
    return sell_fill(ptr_host, col_host, val_host, num_rows, lanes, slice_off_host, slot_row_host, sell_col_host, sell_val_host);
}
int prfdd_sell_fill_f32(const int *ptr_host, const int *col_host, const double *val_host, int num_rows, int lanes, const int *slice_off_host, const int *slot_row_host, int *sell_col_host, float *sell_val_host)
{
    return sell_fill(ptr_host, col_host, val_host, num_rows, lanes, slice_off_host, slot_row_host, sell_col_host, sell_val_host);
}

// ---- descriptor entry points ----------------------------------------------------------------
int prfdd_csrm_multiply(double *Au, const prfdd_csr_matrix *A, const double *u, prfdd_stream_t stream) { return t_multiply(Au, view(*A), u, S(stream)); }
int prfdd_csrm_multiply_range(double *Au, const prfdd_csr_matrix *A, const double *u, int row_start, int row_end, prfdd_stream_t stream) { return t_multiply_range(Au, view(*A), u, row_start, row_end, S(stream)); }
int prfdd_csrm_multiply_weight(double *Au, const prfdd_csr_matrix *A, const double *u, const double *weight, prfdd_stream_t stream) { return t_multiply_weight(Au, view(*A), u, weight, S(stream)); }
int prfdd_csrm_matvec(double *y, const prfdd_csr_matrix *A, const double *x, double alpha, double beta, prfdd_stream_t stream) { return t_matvec(y, view(*A), x, alpha, beta, S(stream)); }
int prfdd_csrm_residual(double *v, const prfdd_csr_matrix *A, const double *u, const double *f, prfdd_stream_t stream) { return t_residual(v, view(*A), u, f, S(stream)); }
int prfdd_csrm_cheby_residual(double *r, double *t, const prfdd_csr_matrix *A, const double *u, const double *f, const double *ds, double c_hi, prfdd_stream_t stream) { return t_cheby_residual(r, t, view(*A), u, f, ds, c_hi, S(stream)); }
int prfdd_csrm_restrict_cheby_residual(double *f, double *r, double *t, const prfdd_csr_matrix *R, const double *v, const double *ds, double c_hi, prfdd_stream_t stream) { return t_restrict_cheby_residual(f, r, t, view(*R), v, ds, c_hi, S(stream)); }
int prfdd_csrm_cheby_step(double *u, double *t_out, const prfdd_csr_matrix *A, const double *t_in, const double *r, const double *ds, double c, int last, int u_is_zero, prfdd_stream_t stream) { return t_cheby_step(u, t_out, view(*A), t_in, r, ds, c, last, u_is_zero, S(stream)); }

// FP32 instances: the V-cycle with `Float float` (AMG/config.hpp:4)
int prfdd_csrm_multiply_f32(float *Au, const prfdd_csr_matrix_f32 *A, const float *u, prfdd_stream_t stream) { return t_multiply(Au, view(*A), u, S(stream)); }
int prfdd_csrm_matvec_f32(float *y, const prfdd_csr_matrix_f32 *A, const float *x, float alpha, float beta, prfdd_stream_t stream) { return t_matvec(y, view(*A), x, alpha, beta, S(stream)); }
int prfdd_csrm_residual_f32(float *v, const prfdd_csr_matrix_f32 *A, const float *u, const float *f, prfdd_stream_t stream) { return t_residual(v, view(*A), u, f, S(stream)); }
int prfdd_csrm_cheby_residual_f32(float *r, float *t, const prfdd_csr_matrix_f32 *A, const float *u, const float *f, const float *ds, float c_hi, prfdd_stream_t stream) { return t_cheby_residual(r, t, view(*A), u, f, ds, c_hi, S(stream)); }
int prfdd_csrm_restrict_cheby_residual_f32(float *f, float *r, float *t, const prfdd_csr_matrix_f32 *R, const float *v, const float *ds, float c_hi, prfdd_stream_t stream) { return t_restrict_cheby_residual(f, r, t, view(*R), v, ds, c_hi, S(stream)); }
int prfdd_csrm_cheby_step_f32(float *u, float *t_out, const prfdd_csr_matrix_f32 *A, const float *t_in, const float *r, const float *ds, float c, int last, int u_is_zero, prfdd_stream_t stream) { return t_cheby_step(u, t_out, view(*A), t_in, r, ds, c, last, u_is_zero, S(stream)); }

int prfdd_dense_solve_f32(float *x, const float *Ainv, const float *b, int n, prfdd_stream_t stream)
{
    if (n <= 0) return 0;
    launch_pdl(k_dense_t<float>, (n + 7) / 8, 256, 0, S(stream), x, Ainv, b, n);
    return launched(4.0 * n * (double)n + 8.0 * n);
}

int prfdd_cast_f64_to_f32(float *dst, const double *src, int n, prfdd_stream_t stream)
{
    if (n <= 0) return 0;
    k_cast<float, double><<<stream_grid(n, 256, 2, 8), 256, 0, S(stream)>>>(dst, src, n);
    return launched(12.0 * n);
}

int prfdd_cast_f32_to_f64(double *dst, const float *src, int n, prfdd_stream_t stream)
{
    if (n <= 0) return 0;
    k_cast<double, float><<<stream_grid(n, 256, 2, 8), 256, 0, S(stream)>>>(dst, src, n);
    return launched(12.0 * n);
}

int prfdd_cheby_order1_f32(float *u, const float *r, const float *ds, float c, int u_is_zero, int size, prfdd_stream_t stream)
{
    // first-order smoother tail in FP32 (the FP64 form lives in k_vector.cu): one launch of the element-wise epilogue, no product
    prfdd_csr_matrix_f32 none = {};
    int dummy = 0;
    none.col = &dummy; none.num_rows = size; none.threads_per_row = 1;
    static const int zero_ptr = 0;
    none.ptr = &zero_ptr; // never read: x == NULL skips the row walk
    if (u_is_zero)
        return spmv(view(none), (const float *)nullptr, 0, size, S(stream), 12.0, [=] __device__(int row, float) { u[row] = ds[row] * (c * r[row]); });
    return spmv(view(none), (const float *)nullptr, 0, size, S(stream), 16.0, [=] __device__(int row, float) { u[row] += ds[row] * (c * r[row]); });
}

// ---- pointer entry points: the reference's kernel arguments + one lanes-per-row hint ------------
int prfdd_csr_multiply(double *Au, const int *ptr, const int *col, const double *val, const double *u, int num_rows, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, num_rows, tpr);
    return prfdd_csrm_multiply(Au, &A, u, stream);
}

int prfdd_csr_multiply_range(double *Au, const int *ptr, const int *col, const double *val, const double *u, int row_start, int row_end, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, row_end + 1, tpr);
    return prfdd_csrm_multiply_range(Au, &A, u, row_start, row_end, stream);
}

int prfdd_csr_multiply_weight(double *Au, const int *ptr, const int *col, const double *val, const double *u, const double *weight, int num_rows, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, num_rows, tpr);
    return prfdd_csrm_multiply_weight(Au, &A, u, weight, stream);
}

int prfdd_csr_matvec(double *y, const int *ptr, const int *col, const double *val, const double *x, double alpha, double beta, int num_rows, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, num_rows, tpr);
    return prfdd_csrm_matvec(y, &A, x, alpha, beta, stream);
}

int prfdd_csr_residual(double *v, const int *ptr, const int *col, const double *val, const double *u, const double *f, int num_rows, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, num_rows, tpr);
    return prfdd_csrm_residual(v, &A, u, f, stream);
}

int prfdd_cheby_residual(double *r, double *t, const int *ptr, const int *col, const double *val, const double *u, const double *f, const double *ds, double c_hi, int num_rows, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, num_rows, tpr);
    return prfdd_csrm_cheby_residual(r, t, &A, u, f, ds, c_hi, stream);
}

int prfdd_restrict_cheby_residual(double *f, double *r, double *t, const int *ptr, const int *col, const double *val, const double *v, const double *ds, double c_hi, int num_rows, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, num_rows, tpr);
    return prfdd_csrm_restrict_cheby_residual(f, r, t, &A, v, ds, c_hi, stream);
}

int prfdd_cheby_step(double *u, double *t_out, const int *ptr, const int *col, const double *val, const double *t_in, const double *r, const double *ds, double c, int last, int u_is_zero, int num_rows, int tpr, prfdd_stream_t stream)
{
    const prfdd_csr_matrix A = plain(ptr, col, val, num_rows, tpr);
    return prfdd_csrm_cheby_step(u, t_out, &A, t_in, r, ds, c, last, u_is_zero, stream);
}

int prfdd_dense_solve(double *x, const double *Ainv, const double *b, int n, prfdd_stream_t stream)
{
    if (n <= 0) return 0;
    launch_pdl(k_dense_solve, (n + 7) / 8, 256, 0, S(stream), x, Ainv, b, n);
    return launched(8.0 * n * (double)n + 16.0 * n);
}

} // extern "C"
