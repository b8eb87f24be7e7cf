// k_sem.cu -- matrix-free SEM Laplacian and polynomial-reduction restriction (sm_100a)
//
// Au = sum_d D_d^T ( G (D u) ) per element, ONE launch, nothing but u, the six geometric factors
// and Au touch HBM (64 B/point in 3D, 40 B/point in 2D).  Replaces the two-launch,
// three-temporary stiffness_matrix_1/2 of domain.okl:5-98 and subdomain.okl:4-101.
//
// 3D layout of the work: one thread per (i,j) column of an element, the k direction is walked in
// registers (u column and Au column live in registers), the (i,j) slices of u / G*Du are exchanged
// through shared memory.  The n x n derivative matrix is held
//   * per thread in registers for the four thread-dependent rows/columns (D[i][.], D[j][.],
//     D[.][i], D[.][j]) when n <= 10,
//   * in the kernel-parameter constant bank for the uniform accesses D[k][m] (compile-time k, m
//     after unrolling => constant-bank operands of DFMA, no load instruction at all).
// Several elements share a CTA so that small degrees still fill warps.
#include "common.cuh"
#include <cstdlib>

namespace prfdd
{
struct G6
{
    const double *g[6];
};

template <int N>
struct DParam
{
    double v[N * N];
};

template <int N>
constexpr int epb3d()
{
    // elements per CTA: aim at 128 threads
    return (N * N >= 128) ? 1 : (128 / (N * N));
}

template <int N, int EPB, int MINB>
__global__ void __launch_bounds__(N *N *EPB, MINB) k_ax3d(double *__restrict__ Au, const double *__restrict__ u, const G6 G, const DParam<N> Dc, const double *__restrict__ Dg, long long first_point, int num_elems)
{
    constexpr int N2 = N * N, N3 = N * N * N;
    constexpr int LD = N | 1; // odd leading dimension: conflict-free row reads by thread-dependent row
    __shared__ double s_u[EPB][N][N];
    __shared__ double s_gr[EPB][N][N];
    __shared__ double s_gs[EPB][N][N];
    __shared__ double s_D[N * LD];

    const int tid = threadIdx.x;
    const int el = tid / N2;
    const int ij = tid - el * N2;
    const int j = ij / N;
    const int i = ij - j * N;
    const long long e = (long long)blockIdx.x * EPB + el;
    const bool active = e < num_elems;
    const long long base = first_point + e * N3 + ij;

    for (int t = tid; t < N2; t += N2 * EPB) s_D[(t / N) * LD + (t % N)] = Dg[t];

    double r_u[N], r_Au[N];
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        r_u[k] = active ? u[base + k * N2] : 0.0;
        r_Au[k] = 0.0;
    }
    __syncthreads();

    // thread-dependent rows / columns of D
    double Di[N], Dj[N], Dti[N], Dtj[N];
#pragma unroll
    for (int m = 0; m < N; m++)
    {
        Di[m] = s_D[i * LD + m];
        Dj[m] = s_D[j * LD + m];
        Dti[m] = s_D[m * LD + i];
        Dtj[m] = s_D[m * LD + j];
    }

    double g_nxt[6];
#pragma unroll
    for (int c = 0; c < 6; c++) g_nxt[c] = active ? G.g[c][base] : 0.0;

#pragma unroll
    for (int k = 0; k < N; k++)
    {
        double g_cur[6];
#pragma unroll
        for (int c = 0; c < 6; c++) g_cur[c] = g_nxt[c];
        if (k + 1 < N)
        {
#pragma unroll
            for (int c = 0; c < 6; c++) g_nxt[c] = active ? G.g[c][base + (k + 1) * N2] : 0.0;
        }

        s_u[el][j][i] = r_u[k];
        __syncthreads();

        double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
        for (int m = 0; m < N; m++)
        {
            ur += Di[m] * s_u[el][j][m];
            us += Dj[m] * s_u[el][m][i];
            ut += Dc.v[k * N + m] * r_u[m];
        }
        const double gr = g_cur[0] * ur + g_cur[3] * us + g_cur[4] * ut;
        const double gs = g_cur[3] * ur + g_cur[1] * us + g_cur[5] * ut;
        const double gt = g_cur[4] * ur + g_cur[5] * us + g_cur[2] * ut;

        s_gr[el][j][i] = gr;
        s_gs[el][j][i] = gs;
        __syncthreads();

        double a = 0.0, b = 0.0;
#pragma unroll
        for (int m = 0; m < N; m++)
        {
            a += Dti[m] * s_gr[el][j][m];
            b += Dtj[m] * s_gs[el][m][i];
        }
        r_Au[k] += a + b;
#pragma unroll
        for (int m = 0; m < N; m++) r_Au[m] += Dc.v[k * N + m] * gt;
    }

    if (active)
    {
#pragma unroll
        for (int k = 0; k < N; k++) Au[base + k * N2] = r_Au[k];
    }
}

// ---------------------------------------------------------------------------------------------
// bulk-async (TMA 1D) variant: ONE element per CTA of N*N threads.  One thread issues seven
// cp.async.bulk copies (u and the six geometric factors of the element, N^3*8 B each, contiguous) that
// complete on an mbarrier; with ~7 CTAs resident per SM the HBM latency of one element's 7*N^3*8 bytes is
// hidden behind the contractions of the others, and no register is spent on prefetching G.  The (i,j)
// slices of u are read straight out of the staged copy (no per-slice store + barrier); G.Du slices are
// double-buffered so one __syncthreads per k suffices.
// Requires N^3*8 and the element's byte offset to be multiples of 16.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int N, int MINB>
__global__ void __launch_bounds__(N *N, MINB) k_ax3d_bulk(double *__restrict__ Au, const double *__restrict__ u, const G6 G, const DParam<N> Dc, const double *__restrict__ Dg, long long first_point, int num_elems)
{
    constexpr int N2 = N * N, N3 = N * N * N;
    constexpr int LD = N | 1; // odd (N + 1 is even for odd N: rows i and i + 2 would share banks; n = 15 measured 171 us against 74 us at n = 16)
    __shared__ __align__(128) double stage[7][N3]; // 0: u, 1..6: G11,G22,G33,G12,G13,G23
    __shared__ __align__(16) double s_gr[2][N2];
    __shared__ __align__(16) double s_gs[2][N2];
    __shared__ double s_D[N * LD];
    __shared__ __align__(8) unsigned long long bar;

    const int ij = threadIdx.x;
    const int j = ij / N;
    const int i = ij - j * N;
    const long long e = blockIdx.x;
    const long long ebase = first_point + e * N3;
    (void)num_elems;

    if (ij == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (ij == 0)
    {
        constexpr uint32_t bytes = N3 * 8;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(7u * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&stage[0][0])), "l"(u + ebase), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
#pragma unroll
        for (int c = 0; c < 6; c++)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&stage[1 + c][0])), "l"(G.g[c] + ebase), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    // overlap with the copies: derivative matrix into shared memory, thread-dependent rows into registers
    for (int t = ij; t < N2; t += N2) s_D[(t / N) * LD + (t % N)] = Dg[t];
    __syncthreads();
    double Di[N], Dj[N], Dti[N], Dtj[N];
#pragma unroll
    for (int m = 0; m < N; m++)
    {
        Di[m] = s_D[i * LD + m];
        Dj[m] = s_D[j * LD + m];
        Dti[m] = s_D[m * LD + i];
        Dtj[m] = s_D[m * LD + j];
    }
    // wait for the bytes
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }

    double r_u[N], r_Au[N];
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        r_u[k] = stage[0][k * N2 + ij];
        r_Au[k] = 0.0;
    }

#pragma unroll
    for (int k = 0; k < N; k++)
    {
        const double *uk = &stage[0][k * N2];
        double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
        for (int m = 0; m < N; m++)
        {
            ur += Di[m] * uk[j * N + m];
            us += Dj[m] * uk[m * N + i];
            ut += Dc.v[k * N + m] * r_u[m];
        }
        const int p = k * N2 + ij;
        const double g0 = stage[1][p], g1 = stage[2][p], g2 = stage[3][p], g3 = stage[4][p], g4 = stage[5][p], g5 = stage[6][p];
        const double gr = g0 * ur + g3 * us + g4 * ut;
        const double gs = g3 * ur + g1 * us + g5 * ut;
        const double gt = g4 * ur + g5 * us + g2 * ut;
        const int b = k & 1;
        s_gr[b][ij] = gr;
        s_gs[b][ij] = gs;
        __syncthreads();
        double a = 0.0, c2 = 0.0;
#pragma unroll
        for (int m = 0; m < N; m++)
        {
            a += Dti[m] * s_gr[b][j * N + m];
            c2 += Dtj[m] * s_gs[b][m * N + i];
        }
        r_Au[k] += a + c2;
#pragma unroll
        for (int m = 0; m < N; m++) r_Au[m] += Dc.v[k * N + m] * gt;
    }

#pragma unroll
    for (int k = 0; k < N; k++) Au[ebase + k * N2 + ij] = r_Au[k];
}

// large degrees (n > 10): D stays in shared memory, everything else identical
// two resident CTAs for n >= 15 (128 registers): n = 16 94.8 -> 74.2 us per 512 elements (measured)
template <int N>
__global__ void __launch_bounds__(N *N, (N >= 15 ? 2 : 1)) k_ax3d_big(double *__restrict__ Au, const double *__restrict__ u, const G6 G, const double *__restrict__ Dg, long long first_point, int num_elems)
{
    constexpr int N2 = N * N, N3 = N * N * N;
    constexpr int LD = N | 1; // odd (N + 1 is even for odd N: rows i and i + 2 would share banks; n = 15 measured 171 us against 74 us at n = 16)
    __shared__ double s_u[N][N];
    __shared__ double s_gr[N][N];
    __shared__ double s_gs[N][N];
    __shared__ double s_D[N * LD];

    const int ij = threadIdx.x;
    const int j = ij / N;
    const int i = ij - j * N;
    const long long e = blockIdx.x;
    const long long base = first_point + e * N3 + ij;
    (void)num_elems;

    s_D[j * LD + i] = Dg[j * N + i];

    double r_u[N], r_Au[N];
#pragma unroll
    for (int k = 0; k < N; k++)
    {
        r_u[k] = u[base + k * N2];
        r_Au[k] = 0.0;
    }
    __syncthreads();

#pragma unroll
    for (int k = 0; k < N; k++)
    {
        double g_cur[6];
#pragma unroll
        for (int c = 0; c < 6; c++) g_cur[c] = G.g[c][base + k * N2];

        s_u[j][i] = r_u[k];
        __syncthreads();

        double ur = 0.0, us = 0.0, ut = 0.0;
#pragma unroll
        for (int m = 0; m < N; m++)
        {
            ur += s_D[i * LD + m] * s_u[j][m];
            us += s_D[j * LD + m] * s_u[m][i];
            ut += s_D[k * LD + m] * r_u[m];
        }
        const double gr = g_cur[0] * ur + g_cur[3] * us + g_cur[4] * ut;
        const double gs = g_cur[3] * ur + g_cur[1] * us + g_cur[5] * ut;
        const double gt = g_cur[4] * ur + g_cur[5] * us + g_cur[2] * ut;

        s_gr[j][i] = gr;
        s_gs[j][i] = gs;
        __syncthreads();

        double a = 0.0, b = 0.0;
#pragma unroll
        for (int m = 0; m < N; m++)
        {
            a += s_D[m * LD + i] * s_gr[j][m];
            b += s_D[m * LD + j] * s_gs[m][i];
        }
#pragma unroll
        for (int m = 0; m < N; m++) r_Au[m] += s_D[k * LD + m] * gt + ((m == k) ? (a + b) : 0.0);
    }

#pragma unroll
    for (int k = 0; k < N; k++) Au[base + k * N2] = r_Au[k];
}

// ---------------------------------------------------------------------------------------------
// n = 16 (degree 15, BASELINE configs[4]) on the FP64 tensor cores.  k_ax3d_big<16> is bound by shared-memory wavefronts (ncu:
// data-pipe wavefronts 76 % of peak, FP64 pipe 22 %, 0.28 of the HBM roofline): every multiply-add of the five contractions
// reads one or two operands from shared memory.  The contractions of one element are 96 products of 16x16 matrices
// (per k plane  Ur = U_k D^T, Us = D U_k;  per j plane  Ut = D T_j;  and the three transposed ones), so here they run as
// mma.sync.m8n8k4.f64: one shared-memory read per operand FRAGMENT (256 multiply-adds) instead of per multiply-add, D and D^T
// fragments in registers.  One element per CTA of 8 warps; warp w owns the k planes 2w, 2w+1 for the r/s products and the
// j planes 2w, 2w+1 for the t products; ut / gt cross between the two ownerships through shared memory (two block barriers),
// G is read once, in accumulator layout (pairs of points), straight from global memory.
// Measured (512 elements, 134 MB): 94.8 us (k_ax3d_big, one CTA per SM) -> 74.2 us (two CTAs) -> 41.8 us = 0.50 of the HBM roofline; a bulk L2
// prefetch of the element's factor blocks (46.9 us) and a register pipeline of the G reads (43.6 us) did not help and are not kept.
//   fragments (lane l, g = l/4, t = l%4):  A(tm,ks) = X[8tm+g][4ks+t],  B(ks,tn) = Y[4ks+t][8tn+g],  C(tm,tn) = Z[8tm+g][8tn+2t+{0,1}]
// Shared layout  i + 20 j + 324 k: both strides are 4 mod 16, so that the 16 lanes of a half warp of every fragment read
// (4 rows x 4 columns) hit 16 different 8-byte banks.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) k_ax3d_mma16(double *__restrict__ Au, const double *__restrict__ u, const G6 G, const double *__restrict__ Dg, long long first_point)
{
    constexpr int N = 16, N2 = 256, N3 = 4096, SJ = 20, SK = 324, PLANE = 16 * SJ;
    extern __shared__ __align__(16) double sm[];
    double *U = sm;                 // u, then per k plane: gr, then the r+s part of Au
    double *W = sm + N * SK;        // ut, then gt
    double *GS = sm + 2 * N * SK;   // per warp: one plane of gs
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long ebase = first_point + (long long)blockIdx.x * N3;
    auto idx = [](int i, int j, int k) { return i + j * SJ + k * SK; };

    // D fragments: DA[tm][ks] = D[8tm+g][4ks+t]  (D as A operand, D^T as B operand);  DB[ks][tn] = D[4ks+t][8tn+g]  (D as B operand, D^T as A operand)
    double DA[2][4], DB[4][2];
#pragma unroll
    for (int ks = 0; ks < 4; ks++)
#pragma unroll
        for (int q = 0; q < 2; q++)
        {
            DA[q][ks] = Dg[(8 * q + g) * N + 4 * ks + t];
            DB[ks][q] = Dg[(4 * ks + t) * N + 8 * q + g];
        }
    // the element's u
#pragma unroll
    for (int r = 0; r < 16; r++)
    {
        const int p = tid + 256 * r;
        U[idx(p & 15, (p >> 4) & 15, r)] = u[ebase + p];
    }
    __syncthreads();

    // t products: ut(.,j,.) = D T_j,  T_j[m][i] = u(i,j,m)
#pragma unroll 1
    for (int jj = 0; jj < 2; jj++)
    {
        const int j = 2 * w + jj;
        double C[2][2][2] = {};
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
#pragma unroll
            for (int tn = 0; tn < 2; tn++)
            {
                const double b = U[idx(8 * tn + g, j, 4 * ks + t)];
                dmma(C[0][tn], DA[0][ks], b);
                dmma(C[1][tn], DA[1][ks], b);
            }
#pragma unroll
        for (int tm = 0; tm < 2; tm++)
#pragma unroll
            for (int tn = 0; tn < 2; tn++) *reinterpret_cast<double2 *>(&W[idx(8 * tn + 2 * t, j, 8 * tm + g)]) = make_double2(C[tm][tn][0], C[tm][tn][1]);
    }
    __syncthreads();

    // r and s products of the warp's k planes, the geometric factors, and the transposed r and s products
    double *gsb = GS + w * PLANE;
#pragma unroll 1
    for (int kk = 0; kk < 2; kk++)
    {
        const int k = 2 * w + kk;
        double CR[2][2][2] = {}, CS[2][2][2] = {};
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
        {
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                const double a = U[idx(4 * ks + t, 8 * q + g, k)]; // U_k[j][m] as A(tm = q)
                dmma(CR[q][0], a, DA[0][ks]);
                dmma(CR[q][1], a, DA[1][ks]);
                const double b = U[idx(8 * q + g, 4 * ks + t, k)]; // U_k[m][i] as B(tn = q)
                dmma(CS[0][q], DA[0][ks], b);
                dmma(CS[1][q], DA[1][ks], b);
            }
        }
        __syncwarp(); // every fragment of U_k has been read: the plane is overwritten below
#pragma unroll
        for (int tm = 0; tm < 2; tm++)
#pragma unroll
            for (int tn = 0; tn < 2; tn++)
            {
                const int i0 = 8 * tn + 2 * t, j = 8 * tm + g;
                const int o = idx(i0, j, k);
                const double2 ut = *reinterpret_cast<const double2 *>(&W[o]);
                const long long p = ebase + i0 + N * j + N2 * k;
                const double2 g0 = *reinterpret_cast<const double2 *>(&G.g[0][p]), g1 = *reinterpret_cast<const double2 *>(&G.g[1][p]),
                              g2 = *reinterpret_cast<const double2 *>(&G.g[2][p]), g3 = *reinterpret_cast<const double2 *>(&G.g[3][p]),
                              g4 = *reinterpret_cast<const double2 *>(&G.g[4][p]), g5 = *reinterpret_cast<const double2 *>(&G.g[5][p]);
                const double ur0 = CR[tm][tn][0], ur1 = CR[tm][tn][1], us0 = CS[tm][tn][0], us1 = CS[tm][tn][1];
                *reinterpret_cast<double2 *>(&U[o]) = make_double2(g0.x * ur0 + g3.x * us0 + g4.x * ut.x, g0.y * ur1 + g3.y * us1 + g4.y * ut.y);
                *reinterpret_cast<double2 *>(&gsb[i0 + j * SJ]) = make_double2(g3.x * ur0 + g1.x * us0 + g5.x * ut.x, g3.y * ur1 + g1.y * us1 + g5.y * ut.y);
                *reinterpret_cast<double2 *>(&W[o]) = make_double2(g4.x * ur0 + g5.x * us0 + g2.x * ut.x, g4.y * ur1 + g5.y * us1 + g2.y * ut.y);
            }
        __syncwarp();
        double CP[2][2][2] = {};
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
        {
#pragma unroll
            for (int q = 0; q < 2; q++)
            {
                const double a = U[idx(4 * ks + t, 8 * q + g, k)]; // Gr_k[j][m] as A(tm = q):  sum_m gr(m,j) D[m][i]
                dmma(CP[q][0], a, DB[ks][0]);
                dmma(CP[q][1], a, DB[ks][1]);
                const double b = gsb[(4 * ks + t) * SJ + 8 * q + g]; // Gs_k[m][i] as B(tn = q):  sum_m D[m][j] gs(i,m)
                dmma(CP[0][q], DB[ks][0], b);
                dmma(CP[1][q], DB[ks][1], b);
            }
        }
        __syncwarp();
#pragma unroll
        for (int tm = 0; tm < 2; tm++)
#pragma unroll
            for (int tn = 0; tn < 2; tn++) *reinterpret_cast<double2 *>(&U[idx(8 * tn + 2 * t, 8 * tm + g, k)]) = make_double2(CP[tm][tn][0], CP[tm][tn][1]);
    }
    __syncthreads();

    // transposed t products of the warp's j planes:  sum_m D[m][k] gt(i,j,m), plus the r+s part, to global memory
#pragma unroll 1
    for (int jj = 0; jj < 2; jj++)
    {
        const int j = 2 * w + jj;
        double C[2][2][2] = {};
#pragma unroll
        for (int ks = 0; ks < 4; ks++)
#pragma unroll
            for (int tn = 0; tn < 2; tn++)
            {
                const double b = W[idx(8 * tn + g, j, 4 * ks + t)];
                dmma(C[0][tn], DB[ks][0], b);
                dmma(C[1][tn], DB[ks][1], b);
            }
#pragma unroll
        for (int tm = 0; tm < 2; tm++)
#pragma unroll
            for (int tn = 0; tn < 2; tn++)
            {
                const int i0 = 8 * tn + 2 * t, k = 8 * tm + g;
                const double2 rs = *reinterpret_cast<const double2 *>(&U[idx(i0, j, k)]);
                *reinterpret_cast<double2 *>(&Au[ebase + i0 + N * j + N2 * k]) = make_double2(C[tm][tn][0] + rs.x, C[tm][tn][1] + rs.y);
            }
    }
}

// 2D: one thread per point
template <int N, int EPB>
__global__ void __launch_bounds__(N *N *EPB) k_ax2d(double *__restrict__ Au, const double *__restrict__ u, const G6 G, const double *__restrict__ Dg, long long first_point, int num_elems)
{
    constexpr int N2 = N * N;
    constexpr int LD = N | 1; // odd (N + 1 is even for odd N: rows i and i + 2 would share banks; n = 15 measured 171 us against 74 us at n = 16)
    __shared__ double s_u[EPB][N][N];
    __shared__ double s_gr[EPB][N][N];
    __shared__ double s_gs[EPB][N][N];
    __shared__ double s_D[N * LD];

    const int tid = threadIdx.x;
    const int el = tid / N2;
    const int ij = tid - el * N2;
    const int j = ij / N;
    const int i = ij - j * N;
    const long long e = (long long)blockIdx.x * EPB + el;
    const bool active = e < num_elems;
    const long long idx = first_point + e * N2 + ij;

    for (int t = tid; t < N2; t += N2 * EPB) s_D[(t / N) * LD + (t % N)] = Dg[t];
    s_u[el][j][i] = active ? u[idx] : 0.0;
    const double g0 = active ? G.g[0][idx] : 0.0, g1 = active ? G.g[1][idx] : 0.0, g2 = active ? G.g[2][idx] : 0.0;
    __syncthreads();

    double ur = 0.0, us = 0.0;
#pragma unroll
    for (int m = 0; m < N; m++)
    {
        ur += s_D[i * LD + m] * s_u[el][j][m];
        us += s_D[j * LD + m] * s_u[el][m][i];
    }
    s_gr[el][j][i] = g0 * ur + g2 * us;
    s_gs[el][j][i] = g2 * ur + g1 * us;
    __syncthreads();

    double a = 0.0, b = 0.0;
#pragma unroll
    for (int m = 0; m < N; m++)
    {
        a += s_D[m * LD + i] * s_gr[el][j][m];
        b += s_D[m * LD + j] * s_gs[el][m][i];
    }
    if (active) Au[idx] = a + b;
}

template <int N>
static int launch_ax3d(double *Au, const double *u, const G6 &G, const double *D_host, const double *D_dev, long long first_point, int num_elems, cudaStream_t st)
{
    if (num_elems <= 0) return 0;
    if (N <= 10 && !D_host)
    {
        // no host copy of D to put into the constant bank: the generic kernel reads D from device memory
        k_ax3d_big<N><<<num_elems, N * N, 0, st>>>(Au, u, G, D_dev, first_point, num_elems);
        return launched(64.0 * num_elems * N * N * N);
    }
    if constexpr (N <= 10)
    {
        constexpr int EPB = epb3d<N>();
        DParam<N> Dc;
        for (int t = 0; t < N * N; t++) Dc.v[t] = D_host[t];
        if constexpr (N == 6 || N == 8)
        {
            // bulk-async variant: needs 16-byte aligned element blocks (static shared memory: n <= 8)
            static const bool no_bulk = getenv("PRFDD_AX_NO_BULK") != nullptr;
            uintptr_t align = reinterpret_cast<uintptr_t>(u);
            for (int c = 0; c < 6; c++) align |= reinterpret_cast<uintptr_t>(G.g[c]);
            if (!no_bulk && first_point % 2 == 0 && align % 16 == 0)
            {
                constexpr int MINB = (N == 6) ? 10 : 6;
                k_ax3d_bulk<N, MINB><<<num_elems, N * N, 0, st>>>(Au, u, G, Dc, D_dev, first_point, num_elems);
                return launched(64.0 * num_elems * N * N * N);
            }
        }
        int grid = (num_elems + EPB - 1) / EPB;
        k_ax3d<N, EPB, 1><<<grid, N * N * EPB, 0, st>>>(Au, u, G, Dc, D_dev, first_point, num_elems);
    }
    else
    {
        if constexpr (N == 16)
        {
            // tensor-core variant: pairs of points are read and written as 16-byte accesses
            static const bool no_mma = getenv("PRFDD_AX_NO_MMA") != nullptr;
            uintptr_t align = reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(Au);
            for (int c = 0; c < 6; c++) align |= reinterpret_cast<uintptr_t>(G.g[c]);
            if (!no_mma && first_point % 2 == 0 && align % 16 == 0)
            {
                constexpr int smem = (2 * 16 * 324 + 8 * 16 * 20) * (int)sizeof(double);
                static const cudaError_t attr = cudaFuncSetAttribute(k_ax3d_mma16, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                if (attr != cudaSuccess) return (int)attr;
                k_ax3d_mma16<<<num_elems, 256, smem, st>>>(Au, u, G, D_dev, first_point);
                return launched(64.0 * num_elems * N * N * N);
            }
        }
        k_ax3d_big<N><<<num_elems, N * N, 0, st>>>(Au, u, G, D_dev, first_point, num_elems);
    }
    return launched(64.0 * num_elems * N * N * N); // u 8 + six factors 48 + Au 8 per point
}

template <int N>
static int launch_ax2d(double *Au, const double *u, const G6 &G, const double *D_dev, long long first_point, int num_elems, cudaStream_t st)
{
    if (num_elems <= 0) return 0;
    constexpr int EPB = (N * N >= 128) ? 1 : (128 / (N * N));
    int grid = (num_elems + EPB - 1) / EPB;
    k_ax2d<N, EPB><<<grid, N * N * EPB, 0, st>>>(Au, u, G, D_dev, first_point, num_elems);
    return launched(40.0 * num_elems * N * N); // u 8 + three factors 24 + Au 8 per point
}

// D_host: host copy of D_dev (same n*n values) or NULL.  The 3D kernels for n <= 10 take D as a kernel parameter (constant bank);
// without a host copy the generic kernel, which reads D from device memory, is used -- nothing is cached by address.
static int ax_dispatch(double *Au, const double *u, const G6 &G, const double *D_dev, const double *D_host, long long first_point, int num_elems, int n, int dim, cudaStream_t st)
{
    if (num_elems <= 0) return 0;
    if (dim == 2)
    {
        switch (n)
        {
#define C2(N) case N: return launch_ax2d<N>(Au, u, G, D_dev, first_point, num_elems, st);
            C2(2) C2(3) C2(4) C2(5) C2(6) C2(7) C2(8) C2(9) C2(10) C2(11) C2(12) C2(13) C2(14) C2(15) C2(16)
#undef C2
        default: return -4;
        }
    }
    const double *Dh = D_host;
    switch (n)
    {
#define C3(N) case N: return launch_ax3d<N>(Au, u, G, Dh, D_dev, first_point, num_elems, st);
        C3(2) C3(3) C3(4) C3(5) C3(6) C3(7) C3(8) C3(9) C3(10) C3(11) C3(12) C3(13) C3(14) C3(15) C3(16)
#undef C3
    default: return -4;
    }
}

// ---------------------------------------------------------------------------------------------
// restriction: u_c = (J^T (x) J^T (x) J^T) u_f, all directions fused, one element per CTA
// order of the contractions and of each sum follows subdomain.okl:284-366
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_restrict3d(double *__restrict__ uc, const double *__restrict__ J, const double *__restrict__ uf, int nf, int nc)
{
    extern __shared__ double sm[];
    double *sJ = sm;                      // nf*nc
    double *a0 = sJ + nf * nc;            // nf^3
    double *a1 = a0 + nf * nf * nf;       // nc*nf*nf
    double *a2 = a1 + nc * nf * nf;       // nc*nc*nf
    const long long e = blockIdx.x;
    const int nf3 = nf * nf * nf, nc3 = nc * nc * nc;
    for (int t = threadIdx.x; t < nf * nc; t += blockDim.x) sJ[t] = J[t];
    for (int t = threadIdx.x; t < nf3; t += blockDim.x) a0[t] = uf[e * nf3 + t];
    __syncthreads();
    // x: a1(i,j,k) = sum_l J[l][i] a0(l,j,k), layout i + j*nc + k*nc*nf
    for (int t = threadIdx.x; t < nc * nf * nf; t += blockDim.x)
    {
        int i = t % nc, j = (t / nc) % nf, k = t / (nc * nf);
        double s = 0.0;
        for (int l = 0; l < nf; l++) s += sJ[i + l * nc] * a0[l + j * nf + k * nf * nf];
        a1[t] = s;
    }
    __syncthreads();
    // y: a2(i,j,k) = sum_l J[l][j] a1(i,l,k), layout i + j*nc + k*nc*nc
    for (int t = threadIdx.x; t < nc * nc * nf; t += blockDim.x)
    {
        int i = t % nc, j = (t / nc) % nc, k = t / (nc * nc);
        double s = 0.0;
        for (int l = 0; l < nf; l++) s += sJ[j + l * nc] * a1[i + l * nc + k * nc * nf];
        a2[t] = s;
    }
    __syncthreads();
    // z
    for (int t = threadIdx.x; t < nc3; t += blockDim.x)
    {
        int i = t % nc, j = (t / nc) % nc, k = t / (nc * nc);
        double s = 0.0;
        for (int l = 0; l < nf; l++) s += sJ[k + l * nc] * a2[i + j * nc + l * nc * nc];
        uc[e * nc3 + t] = s;
    }
}

__global__ void __launch_bounds__(256) k_restrict2d(double *__restrict__ uc, const double *__restrict__ J, const double *__restrict__ uf, int nf, int nc, int num_elems)
{
    extern __shared__ double sm[];
    double *sJ = sm;            // nf*nc
    double *a0 = sJ + nf * nc;  // nf*nf
    double *a1 = a0 + nf * nf;  // nf*nc  (i fine, j coarse), layout i + j*nf
    const long long e = blockIdx.x;
    (void)num_elems;
    for (int t = threadIdx.x; t < nf * nc; t += blockDim.x) sJ[t] = J[t];
    for (int t = threadIdx.x; t < nf * nf; t += blockDim.x) a0[t] = uf[e * nf * nf + t];
    __syncthreads();
    for (int t = threadIdx.x; t < nf * nc; t += blockDim.x)
    {
        int i = t % nf, j = t / nf;
        double s = 0.0;
        for (int k = 0; k < nf; k++) s += sJ[j + k * nc] * a0[i + k * nf];
        a1[t] = s;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < nc * nc; t += blockDim.x)
    {
        int i = t % nc, j = t / nc;
        double s = 0.0;
        for (int k = 0; k < nf; k++) s += a1[j * nf + k] * sJ[k * nc + i];
        uc[e * nc * nc + t] = s;
    }
}
} // namespace prfdd

using namespace prfdd;

extern "C" {

int prfdd_stiffness_matrix_hd(double *Au, const double *u, const double *D_hat, const double *D_hat_host, const double *const g[6], int num_elements, int n, int dim, prfdd_stream_t stream)
{
    G6 G;
    for (int c = 0; c < 6; c++) G.g[c] = g[c];
    return ax_dispatch(Au, u, G, D_hat, D_hat_host, 0, num_elements, n, dim, S(stream));
}

int prfdd_stiffness_matrix(double *Au, const double *u, const double *D_hat, const double *const g[6], int num_elements, int n, int dim, prfdd_stream_t stream)
{
    return prfdd_stiffness_matrix_hd(Au, u, D_hat, nullptr, g, num_elements, n, dim, stream);
}

int prfdd_stiffness_matrix_region_hd(double *Au, const double *u, const double *const g[6], int num_buckets, const int *first_point, const int *num_elements, const int *n, const double *const *D_hat, const double *const *D_hat_host, int dim, prfdd_stream_t stream)
{
    G6 G;
    for (int c = 0; c < 6; c++) G.g[c] = g[c];
    for (int b = 0; b < num_buckets; b++)
    {
        int rc = ax_dispatch(Au, u, G, D_hat[b], D_hat_host ? D_hat_host[b] : nullptr, first_point[b], num_elements[b], n[b], dim, S(stream));
        if (rc) return rc;
    }
    return 0;
}

int prfdd_stiffness_matrix_region(double *Au, const double *u, const double *const g[6], int num_buckets, const int *first_point, const int *num_elements, const int *n, const double *const *D_hat, int dim, prfdd_stream_t stream)
{
    return prfdd_stiffness_matrix_region_hd(Au, u, g, num_buckets, first_point, num_elements, n, D_hat, nullptr, dim, stream);
}

int prfdd_restriction(double *u_c, const double *J_cf, const double *u_f, int num_elements, int n_f, int n_c, int dim, prfdd_stream_t stream)
{
    if (num_elements <= 0) return 0;
    if (dim == 3)
    {
        size_t smem = sizeof(double) * ((size_t)n_f * n_c + (size_t)n_f * n_f * n_f + (size_t)n_c * n_f * n_f + (size_t)n_c * n_c * n_f);
        if (smem > 48 * 1024)
        {
            cudaError_t e = cudaFuncSetAttribute(k_restrict3d, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
        }
        k_restrict3d<<<num_elements, 256, smem, S(stream)>>>(u_c, J_cf, u_f, n_f, n_c);
    }
    else
    {
        size_t smem = sizeof(double) * ((size_t)n_f * n_c * 2 + (size_t)n_f * n_f);
        k_restrict2d<<<num_elements, 256, smem, S(stream)>>>(u_c, J_cf, u_f, n_f, n_c, num_elements);
    }
    const double pf = dim == 3 ? (double)n_f * n_f * n_f : (double)n_f * n_f, pc = dim == 3 ? (double)n_c * n_c * n_c : (double)n_c * n_c;
    return launched(8.0 * num_elements * (pf + pc));
}

} // extern "C"
