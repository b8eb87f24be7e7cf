/*
 * prfdd_b200.h -- C ABI of libprfdd_b200.so
 *
 * B200-native (sm_100a) implementation of the PR-FDD preconditioned Krylov hot path of
 * metalcycling/polynomial_reduction_with_full_domain_decomposition_preconditioner.
 *
 * The reference has no FFI layer; its seam is the OCCA kernel call (one functor per OKL @kernel) plus
 * the five `extern "C"` launchers of AMG/kernels.cu.  This header keeps that convention: plain
 * device pointers, sizes, scalars by value and a cudaStream_t (passed as void*), no torch / C++
 * types.  Every entry point names the reference interface it replaces (file:line in the reference
 * tree).  All functions return 0 on success or a cudaError_t / negative library code, and are
 * asynchronous on `stream` unless stated otherwise.
 *
 * There is NO CPU fallback behind any of these symbols.
 *
 * Conventions
 *   - all floating-point data is FP64, index arrays are int32, global ids int64
 *   - element-local layout: point (i,j,k) of element e lives at e*n^dim + i + j*n + k*n*n
 *   - geometric factors: six separate arrays g[0..5] = [G11,G22,G33,G12,G13,G23]
 *     (2D: g[0],g[1],g[2] = G11,G22,G12)                                   domain.okl:29-30, 47-49
 *   - reductions are deterministic (fixed partition, fixed summation order) and leave their
 *     result in DEVICE memory; nothing synchronises with the host
 */
#ifndef PRFDD_B200_H
#define PRFDD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef void *prfdd_stream_t;    /* cudaStream_t */
typedef struct prfdd_reduce_ws prfdd_reduce_ws; /* scratch for deterministic device reductions */

/* ---------------------------------------------------------------------------------------------
 * library / device
 * ------------------------------------------------------------------------------------------- */
const char *prfdd_version(void);
const char *prfdd_error_string(int code);
int prfdd_device_count(void);
/* replaces occa device.setup({mode CUDA, device_id}) poisson.cpp:136-138 */
int prfdd_set_device(int device);
/* number of kernel launches issued by this library since the last reset (bench.py "gpu_launches") */
long long prfdd_launch_count(void);
void prfdd_launch_count_reset(void);
/* kernels replayed from a captured CUDA graph are counted through this */
void prfdd_launch_count_add(long long n);
/* ALGORITHMIC bytes of the kernels launched since the last reset: every launcher adds what its kernel must move at least
 * (each operand once: SURVEY 8d; e.g. 64 B/point for the SEM operator, 12*nnz + 4*(rows+1) + 8*cols + epilogue operands for an
 * SpMV).  bench.py divides the sum over a solve by the solve time for its whole-solve roofline fraction. */
double prfdd_algorithmic_bytes(void);
void prfdd_algorithmic_bytes_reset(void);
void prfdd_algorithmic_bytes_add(double bytes);

/* replaces occa::memory malloc / free / copyFrom / copyTo / device.finish */
int prfdd_malloc(void **dptr, size_t bytes);
int prfdd_free(void *dptr);
int prfdd_memcpy_h2d(void *dst, const void *src, size_t bytes, prfdd_stream_t stream);
int prfdd_memcpy_d2h(void *dst, const void *src, size_t bytes, prfdd_stream_t stream);
int prfdd_memcpy_d2d(void *dst, const void *src, size_t bytes, prfdd_stream_t stream);
int prfdd_stream_synchronize(prfdd_stream_t stream);

int prfdd_reduce_ws_create(prfdd_reduce_ws **ws);
int prfdd_reduce_ws_destroy(prfdd_reduce_ws *ws);

/* ---------------------------------------------------------------------------------------------
 * matrix-free SEM Laplacian
 * ------------------------------------------------------------------------------------------- */
/* Au = sum_d D_d^T (G (D u)) on `num_elements` elements of degree n-1, fused in one launch.
 * replaces stiffness_matrix_1 + stiffness_matrix_2, domain.okl:5-98 (Domain::stiffness_matrix,
 * domain.tpp:602-609).  D_hat is row-major n*n, D_hat[i*n+m] = dl_m/dxi(xi_i) (domain.tpp:312). */
int prfdd_stiffness_matrix(double *Au, const double *u, const double *D_hat, const double *const g[6],
                           int num_elements, int n, int dim, prfdd_stream_t stream);
/* the same with a HOST copy of D_hat (same n*n values).  The fastest 3D kernels (n <= 10) take D as a kernel parameter, i.e. from
 * the constant bank, which can only be filled from host memory; without the host copy (the form above) the generic kernel that
 * reads D from device memory runs instead.  Nothing is cached by address and no entry point synchronises, so both forms are
 * safe under stream capture and with reused allocations.  The bulk-async variant (n = 6, 8) additionally needs u and g[0..5]
 * 16-byte aligned; otherwise the register-staged kernel runs. */
int prfdd_stiffness_matrix_hd(double *Au, const double *u, const double *D_hat, const double *D_hat_host, const double *const g[6],
                              int num_elements, int n, int dim, prfdd_stream_t stream);

/* variable-degree composite operator: the region is a sequence of `num_buckets` contiguous runs of
 * equal-degree elements (run b: first point first_point[b], num_elements[b], n[b] points per
 * direction, derivative matrix D_hat[b]).  replaces stiffness_matrix_1/2 of subdomain.okl:4-101
 * (Subdomain::stiffness_matrix, subdomain.tpp:3953-3966) whose per-point offset/vert/level lookups
 * become per-run launch parameters. */
int prfdd_stiffness_matrix_region(double *Au, const double *u, const double *const g[6], int num_buckets,
                                  const int *first_point, const int *num_elements, const int *n,
                                  const double *const *D_hat, int dim, prfdd_stream_t stream);
/* D_hat_host[b]: host copy of D_hat[b] (see prfdd_stiffness_matrix_hd); NULL = none */
int prfdd_stiffness_matrix_region_hd(double *Au, const double *u, const double *const g[6], int num_buckets,
                                     const int *first_point, const int *num_elements, const int *n,
                                     const double *const *D_hat, const double *const *D_hat_host, int dim, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * gather-scatter (direct stiffness summation), index-map form of Q / Q^T
 * ------------------------------------------------------------------------------------------- */
/* nodes[v] = (weight ? weight[v] : 1) * sum_{j in [ptr[v],ptr[v+1])} u[col[j]]
 * replaces CSR_Matrix::multiply / multiply_weight on Qt, csr_matrix.okl:5-18, 35-48 (all values
 * of Q^T are 1.0, domain.tpp:289-294) */
int prfdd_gather(double *nodes, const int *ptr, const int *col, const double *u, const double *weight,
                 int num_nodes, prfdd_stream_t stream);
/* out[p] = (mask ? mask[p] : 1) * nodes[node_of_point[p]]
 * replaces CSR_Matrix::multiply / multiply_weight on Q (one 1.0 per row), domain.tpp:596-599 */
int prfdd_scatter(double *out, const int *node_of_point, const double *nodes, const double *mask,
                  int num_points, prfdd_stream_t stream);
/* halo: buf[i] = nodes[idx[i]]   /   nodes[idx[i]] += buf[i]   (replaces the D2H + gslib_gs + H2D of
 * domain.tpp:590-594; idx lists are strictly increasing per peer so the add needs no atomics) */
int prfdd_halo_pack(double *buf, const double *nodes, const int *idx, int count, prfdd_stream_t stream);
int prfdd_halo_unpack_add(double *nodes, const double *buf, const int *idx, int count, prfdd_stream_t stream);
/* dst[idx[i]] = buf[i]: unpack of the tree exchange into the region slots (replaces gslib_gs + H2D, subdomain.tpp:4625-4631) */
int prfdd_scatter_assign(double *dst, const double *buf, const int *idx, int count, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * CSR SpMV (csr_matrix.okl, AMG/csr_matrix.cpp:126-133 cusparseSpMV)
 * ------------------------------------------------------------------------------------------- */
/* every SpMV takes one launch-shape hint after the reference's argument list: threads_per_row in
 * {1,2,4,8,16,32} (a sub-warp of that many lanes walks one row with coalesced col/val reads and a
 * shuffle reduction); 0 = library default.  Callers pick it from nnz/rows of the matrix.
 * Row-indexed arguments may be offset to run a range of rows with its own hint (ptr + r0, outputs + r0; col, val and the
 * gathered vector unshifted). */
/* Matrix descriptor: the CSR arrays plus the launch plan of the matrix.  It replaces the process-global registration of
 * round 1 (a list keyed by the device address of `ptr`): everything a launch needs travels in the call.
 *   val == NULL   every stored value is 1.0 (Q and Q^T of a conforming region, domain.tpp:286-294): the value stream is
 *                 never read;
 *   ptr == NULL   exactly one entry per row, row i holds entry i (Q of a conforming region): index-map kernel;
 *   long_rows     device list of the rows longer than long_row_threshold entries: the row kernels skip them and a second
 *                 launch gives each a whole warp (hanging-node rows of the composite grid); optional.
 * prfdd_csr_plan fills the plan from the HOST copy of ptr.  A descriptor is plain data: copy it freely, keep the device
 * arrays alive while it is in use. */
typedef struct prfdd_csr_matrix
{
    const int *ptr;
    const int *col;
    const double *val;
    int num_rows;
    int num_cols;   /* 0: unknown (only used for the algorithmic byte count) */
    int num_nnz;
    int threads_per_row;
    const int *long_rows;
    int num_long_rows;
    int long_row_threshold;
    /* optional sliced copy of the same matrix (prfdd_sell_layout / prfdd_sell_fill): when sell_col is set, full products
     * (all rows, x given) read it instead of ptr/col/val; the arithmetic and its order are those of the row kernels */
    const int *sell_off;    /* [sell_num_slices + 1] first entry of each slice, multiples of 32 */
    const int *sell_col;    /* [sell_off[sell_num_slices]] */
    const double *sell_val;
    const int *sell_row;    /* [sell_num_slices * 32 / sell_lanes] row of each slot, -1: none; NULL: slot = row.  Rows on the
                             * long_rows list must NOT be in the sliced copy (slot row -1): they get their warp-per-row launch */
    int sell_num_slices;
    int sell_lanes;         /* lanes per row: 1, 2, 4, 8, 16 or 32 */
    int sell_window;        /* window_rows the layout was sorted with (0: row order); 256 selects the CTA-per-window kernel */
} prfdd_csr_matrix;
/* host-side planning: sets threads_per_row and long_row_threshold of *A from
 * ptr_host[0..num_rows] (A->num_rows must be set; A->num_nnz is set to ptr_host[num_rows]).  The rows longer than the
 * threshold are written to long_rows_host (capacity entries) and their count is returned (0: no list needed; the caller
 * uploads the list and sets A->long_rows / A->num_long_rows).  Returns -(count) if the capacity is too small. */
int prfdd_csr_plan(prfdd_csr_matrix *A, const int *ptr_host, int *long_rows_host, int capacity);
/* Sliced layout (SELL-C-sigma with `lanes` lanes per row): 32/lanes consecutive slots form a slice; entry k*lanes + t of the
 * row in slot q of a slice sits at sell_off[slice] + 32*k + q*lanes + t, so one warp load of col (or val) is one contiguous
 * 128 (256) byte line whatever the row lengths -- the row kernels' strided segments cost an L1 tag look-up per segment,
 * which is what bounds them (profiles/r2_notes.txt).  Slices are padded to their longest row (value 0); rows are sorted by
 * length inside windows of window_rows consecutive rows (0: keep the order) to keep the padding small.
 * prfdd_sell_layout: fills slice_off_host[num_slices + 1] and slot_row_host[num_slices * 32 / lanes] (-1: empty slot),
 * num_slices = ceil(num_rows / (32/lanes)); returns the padded entry count, or < 0.
 * prfdd_sell_fill_*: writes the padded col/val arrays of that layout. */
long long prfdd_sell_layout(const int *ptr_host, int num_rows, int lanes, int window_rows, int *slice_off_host, int *slot_row_host);
int prfdd_sell_fill(const int *ptr_host, const int *col_host, const double *val_host, int num_rows, int lanes,
                    const int *slice_off_host, const int *slot_row_host, int *sell_col_host, double *sell_val_host);
int prfdd_sell_fill_f32(const int *ptr_host, const int *col_host, const double *val_host, int num_rows, int lanes,
                        const int *slice_off_host, const int *slot_row_host, int *sell_col_host, float *sell_val_host);
/* descriptor forms of the entry points below (same arithmetic, same epilogues) */
int prfdd_csrm_multiply(double *Au, const prfdd_csr_matrix *A, const double *u, prfdd_stream_t stream);
int prfdd_csrm_multiply_range(double *Au, const prfdd_csr_matrix *A, const double *u, int row_start, int row_end, prfdd_stream_t stream);
int prfdd_csrm_multiply_weight(double *Au, const prfdd_csr_matrix *A, const double *u, const double *weight, prfdd_stream_t stream);
int prfdd_csrm_matvec(double *y, const prfdd_csr_matrix *A, const double *x, double alpha, double beta, prfdd_stream_t stream);
int prfdd_csrm_residual(double *v, const prfdd_csr_matrix *A, const double *u, const double *f, prfdd_stream_t stream);
int prfdd_csrm_cheby_residual(double *r, double *t, const prfdd_csr_matrix *A, const double *u, const double *f, const double *ds,
                              double c_hi, prfdd_stream_t stream);
int prfdd_csrm_restrict_cheby_residual(double *f, double *r, double *t, const prfdd_csr_matrix *R, const double *v, const double *ds,
                                       double c_hi, prfdd_stream_t stream);
int prfdd_csrm_cheby_step(double *u, double *t_out, const prfdd_csr_matrix *A, const double *t_in, const double *r, const double *ds,
                          double c, int last, int u_is_zero, prfdd_stream_t stream);
/* y = A x                                        CSR_Matrix::multiply       csr_matrix.okl:5-18 */
int prfdd_csr_multiply(double *Au, const int *ptr, const int *col, const double *val, const double *u,
                       int num_rows, int threads_per_row, prfdd_stream_t stream);
/* rows [row_start,row_end] inclusive              multiply_range            csr_matrix.okl:20-33 */
int prfdd_csr_multiply_range(double *Au, const int *ptr, const int *col, const double *val, const double *u,
                             int row_start, int row_end, int threads_per_row, prfdd_stream_t stream);
/* y = (A x) .* weight                             multiply_weight           csr_matrix.okl:35-48 */
int prfdd_csr_multiply_weight(double *Au, const int *ptr, const int *col, const double *val, const double *u,
                              const double *weight, int num_rows, int threads_per_row, prfdd_stream_t stream);
/* y = alpha A x + beta y                          amg::CSR_Matrix::matvec   AMG/csr_matrix.cpp:112-134 */
int prfdd_csr_matvec(double *y, const int *ptr, const int *col, const double *val, const double *x,
                     double alpha, double beta, int num_rows, int threads_per_row, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * AMG smoother pieces -- same names and argument order as the reference's own C ABI
 * (AMG/kernels.cu:18, 36, 54, 71, 89) plus the fused forms the V-cycle uses
 * ------------------------------------------------------------------------------------------- */
int prfdd_vector_set_to_value(double *data, double value, int size, prfdd_stream_t stream);
int prfdd_main_scaled_residual(double *Sr, double *w, const double *f_m_Au, const double *S, double alpha,
                               int size, prfdd_stream_t stream);
int prfdd_main_polynomial_evaluation(double *w, double *v, const double *r, const double *D_val, double alpha,
                                     int size, prfdd_stream_t stream);
int prfdd_main_update_field(double *u, const double *w, const double *D_val, int size, prfdd_stream_t stream);
int prfdd_vector_multiplication(double *uv, const double *u, const double *v, int size, prfdd_stream_t stream);
/* fused Chebyshev stage 1:  r = ds.*(f - A u) (A u skipped when u == NULL, i.e. u = 0);
 *                           t = ds.*(c_hi*r)          replaces scaled_residual + vector_multiplication
 *                           (subdomain.tpp:19-40, 64) */
int prfdd_cheby_residual(double *r, double *t, const int *ptr, const int *col, const double *val, const double *u,
                         const double *f, const double *ds, double c_hi, int num_rows, int threads_per_row,
                         prfdd_stream_t stream);
/* restriction fused with the head of the next level's zero-guess smoothing: f = R v, r = ds f, t = ds (c_hi r)
 * (= prfdd_csr_multiply followed by prfdd_cheby_residual with u = NULL on the coarse level; one pass, one launch) */
int prfdd_restrict_cheby_residual(double *f, double *r, double *t, const int *ptr, const int *col, const double *val, const double *v,
                                  const double *ds, double c_hi, int num_rows, int threads_per_row, prfdd_stream_t stream);
/* fused Chebyshev Horner step: w = c*r + ds.*(A t_in); if last: u (+)= ds.*w else t_out = ds.*w
 * replaces polynomial_evaluation (+ update_field), subdomain.tpp:45-83.  u_is_zero: u = ds.*w */
int prfdd_cheby_step(double *u, double *t_out, const int *ptr, const int *col, const double *val, const double *t_in,
                     const double *r, const double *ds, double c, int last, int u_is_zero, int num_rows,
                     int threads_per_row, prfdd_stream_t stream);
/* first-order (cheby_order 1) smoother tail: u (+)= ds.*(c*r) */
int prfdd_cheby_order1(double *u, const double *r, const double *ds, double c, int u_is_zero, int size,
                       prfdd_stream_t stream);
/* v = f - A u                                    subdomain.tpp:3660-3661 */
int prfdd_csr_residual(double *v, const int *ptr, const int *col, const double *val, const double *u, const double *f,
                       int num_rows, int threads_per_row, prfdd_stream_t stream);
/* x = Ainv b for the coarsest level (dense, row-major n*n); replaces hypre_GaussElimSolve, subdomain.tpp:4080-4088 */
int prfdd_dense_solve(double *x, const double *Ainv, const double *b, int n, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * FP32 V-cycle: the reference builds its AMG classes and its Subdomain in `Float`
 * (AMG/config.hpp:4, config.hpp:19-20 PTYPE Float); with `Float float` the hierarchy, the smoother and the
 * cycle vectors are single precision under the FP64 outer solve (the outer residual stays an FP64 quantity).
 * Same kernels as above instantiated for float: float values, float vectors, float row sums.
 * ------------------------------------------------------------------------------------------- */
typedef struct prfdd_csr_matrix_f32
{
    const int *ptr;
    const int *col;
    const float *val;
    int num_rows;
    int num_cols;
    int num_nnz;
    int threads_per_row;
    const int *long_rows;
    int num_long_rows;
    int long_row_threshold;
    const int *sell_off;
    const int *sell_col;
    const float *sell_val;
    const int *sell_row;
    int sell_num_slices;
    int sell_lanes;
    int sell_window;
} prfdd_csr_matrix_f32;
int prfdd_csrm_multiply_f32(float *Au, const prfdd_csr_matrix_f32 *A, const float *u, prfdd_stream_t stream);
int prfdd_csrm_matvec_f32(float *y, const prfdd_csr_matrix_f32 *A, const float *x, float alpha, float beta, prfdd_stream_t stream);
int prfdd_csrm_residual_f32(float *v, const prfdd_csr_matrix_f32 *A, const float *u, const float *f, prfdd_stream_t stream);
int prfdd_csrm_cheby_residual_f32(float *r, float *t, const prfdd_csr_matrix_f32 *A, const float *u, const float *f, const float *ds,
                                  float c_hi, prfdd_stream_t stream);
int prfdd_csrm_restrict_cheby_residual_f32(float *f, float *r, float *t, const prfdd_csr_matrix_f32 *R, const float *v, const float *ds,
                                           float c_hi, prfdd_stream_t stream);
int prfdd_csrm_cheby_step_f32(float *u, float *t_out, const prfdd_csr_matrix_f32 *A, const float *t_in, const float *r, const float *ds,
                              float c, int last, int u_is_zero, prfdd_stream_t stream);
int prfdd_cheby_order1_f32(float *u, const float *r, const float *ds, float c, int u_is_zero, int size, prfdd_stream_t stream);
int prfdd_dense_solve_f32(float *x, const float *Ainv, const float *b, int n, prfdd_stream_t stream);
/* the casts at the cycle's boundary: copy_from_domain_data / copy_to_domain_data with EType != DType (subdomain.okl:268-282) */
int prfdd_cast_f64_to_f32(float *dst, const double *src, int n, prfdd_stream_t stream);
int prfdd_cast_f32_to_f64(double *dst, const float *src, int n, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * polynomial reduction (restriction along the degree ladder)
 * ------------------------------------------------------------------------------------------- */
/* u_c = (J^T x J^T [x J^T]) u_f per element, all directions fused.  J is n_f x n_c row-major,
 * J[i*n_c+j] = h_j^coarse(xi_i^fine) (subdomain.tpp:157-159).
 * replaces restriction_1/2/3, subdomain.okl:284-366 (Subdomain::tree_operator, subdomain.tpp:4576-4609) */
int prfdd_restriction(double *u_c, const double *J_cf, const double *u_f, int num_elements, int n_f, int n_c,
                      int dim, prfdd_stream_t stream);
/* precision casts of subdomain.okl:268-282 (EType <-> DType; both double here) */
int prfdd_copy_from_domain_data(double *u, const double *v, int num_points, prfdd_stream_t stream);
int prfdd_copy_to_domain_data(double *u, const double *v, int num_points, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * vector algebra: math.okl:5-35, domain.okl:100-107, 186-193, 226-233
 * ------------------------------------------------------------------------------------------- */
int prfdd_set_to_value(double *u, double alpha, int n, int offset, prfdd_stream_t stream);
int prfdd_invert_vector_elements(double *u, int n, prfdd_stream_t stream);
int prfdd_vector_vector_addition(double *uv, double alpha, const double *u, double beta, const double *v, int n,
                                 prfdd_stream_t stream);
int prfdd_vector_scaling(double *au, double alpha, const double *u, int n, prfdd_stream_t stream);
int prfdd_initialize_arrays(double *u_k, double *r_k, const double *f, int n, prfdd_stream_t stream);
/* u += alpha p ; r1 = r - alpha q.  alpha by value (reference signature) ... */
int prfdd_solution_and_residual_update(double *u_k, double *r_kp1, const double *r_k, const double *p_k,
                                       const double *q_k, double alpha_k, int n, prfdd_stream_t stream);
/* ... or alpha = num[0]/den[0] read from device memory (no host round trip) */
int prfdd_solution_and_residual_update_dev(double *u_k, double *r_kp1, const double *r_k, const double *p_k,
                                           const double *q_k, const double *num, const double *den, int n,
                                           prfdd_stream_t stream);
/* p = z + beta p ; r = r1 */
int prfdd_residual_and_search_update(double *p_k, double *r_k, const double *z_k, const double *r_kp1, double beta_k,
                                     int n, prfdd_stream_t stream);
int prfdd_residual_and_search_update_dev(double *p_k, double *r_k, const double *z_k, const double *r_kp1,
                                         const double *num, const double *den, int n, prfdd_stream_t stream);
/* device-scalar forms used by the Krylov drivers: out = a*x (a = scale_num/scale_den on device) and
 * y += sum_i coef[i*stride] * X_i  (coef on device, sign applied), sequential in i like the
 * reference's chain of vector_vector_addition calls (domain.tpp:817-822, subdomain.tpp:4396-4401) */
int prfdd_vector_scaling_dev(double *au, const double *num, const double *den, const double *u, int n,
                             prfdd_stream_t stream);
int prfdd_multi_axpy_dev(double *y, const double *const *X, const double *coef, int coef_stride, double sign,
                         int count, int n, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * reductions: domain.okl:109-184, 195-264; subdomain.okl:103-209, 229-258.  `out` is DEVICE memory.
 * The reference writes P/128 block partials, copies them to the host and sums serially
 * (domain.tpp:922-926); here the sum is finished on the device by the last block, in a fixed order.
 * ------------------------------------------------------------------------------------------- */
/* out[0] = sum r*qqr*mask                          residual_norm             domain.okl:109-138 */
int prfdd_residual_norm(prfdd_reduce_ws *ws, double *out, const double *r_k, const double *QQt_r_k,
                        const double *dirichlet_mask, int n, prfdd_stream_t stream);
/* out[0] = sum z*r ; out[1] = sum p*q              projection_inner_products domain.okl:140-184 */
int prfdd_projection_inner_products(prfdd_reduce_ws *ws, double *out, const double *z_k, const double *r_k,
                                    const double *p_k, const double *q_k, int n, prfdd_stream_t stream);
/* out[0] = sum (r1-r)*z                            inner_product_flexible    domain.okl:195-224 */
int prfdd_inner_product_flexible(prfdd_reduce_ws *ws, double *out, const double *r_k, const double *r_kp1,
                                 const double *z_k, int n, prfdd_stream_t stream);
/* out[0] = sum u*v*mask                            inner_product             domain.okl:235-264 */
int prfdd_inner_product(prfdd_reduce_ws *ws, double *out, const double *u_k, const double *v_k,
                        const double *dirichlet_mask, int n, prfdd_stream_t stream);
/* out[0] = sum u*v*w (w may be NULL)               (weighted_)inner_product  subdomain.okl:103-163 */
int prfdd_weighted_inner_product(prfdd_reduce_ws *ws, double *out, const double *u, const double *v, const double *w,
                                 int n, prfdd_stream_t stream);
/* out[i] = sum u*V_i*w for i < count (<= 32)       the j+1 Gram-Schmidt dots of one Arnoldi column in one
 *                                                  pass (subdomain.tpp:4389-4394, domain.tpp:810-815) */
int prfdd_multi_inner_product(prfdd_reduce_ws *ws, double *out, const double *u, const double *const *V,
                              const double *w, int count, int n, prfdd_stream_t stream);
/* fused Arnoldi column of the rank-local GMRES (Subdomain::generalized_minimum_residual, subdomain.tpp:4389-4458), working on the
 * assembled copies aV_k = Q^T V_k that are kept from the previous steps (Q^T is linear):
 *   prfdd_orthogonalize_norm   aq <- aq - sum_k coef[k] aV_k ;  out[0] = sum_i w[i] aq[i]^2
 *   prfdd_arnoldi_next         V_next = scale (q - sum_k coef[k] V_k)  (n entries);  aV_next = scale aq  (n_assembled entries)
 * coef, scale and out live in device memory; count <= 32.  They replace the reference's j+1 vector_vector_addition launches, the
 * second assembly + residual_norm of the orthogonalised column and the two vector_scaling launches. */
int prfdd_orthogonalize_norm(prfdd_reduce_ws *ws, double *out, double *aq, const double *const *aV, const double *coef, const double *w,
                             int count, int n, prfdd_stream_t stream);
int prfdd_arnoldi_next(double *V_next, const double *q, const double *const *V, const double *coef, int count, const double *scale, int n,
                       double *aV_next, const double *aq, int n_assembled, prfdd_stream_t stream);
/* out[0] = sum z*r*w ; out[1] = sum p*q*w          projection_inner_products subdomain.okl:165-209 */
int prfdd_weighted_projection_inner_products(prfdd_reduce_ws *ws, double *out, const double *z_k, const double *r_k,
                                             const double *p_k, const double *q_k, const double *weight, int n,
                                             prfdd_stream_t stream);
/* out[0] = sum (r1-r)*z*w                          search_update_inner_product subdomain.okl:229-258 */
int prfdd_search_update_inner_product(prfdd_reduce_ws *ws, double *out, const double *r_k, const double *r_kp1,
                                      const double *z_k, const double *weight, int n, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * device-resident Krylov bookkeeping: the Hessenberg / Givens / back-substitution scalars of the
 * inner flexible GMRES (subdomain.tpp:4404-4470) and the alpha/beta of the inner flexible CG
 * (subdomain.tpp:4215, 4256) live in this struct IN DEVICE MEMORY and are updated by one-thread
 * kernels, so an entire preconditioner application runs without a host synchronisation and can be
 * captured in a CUDA graph.  Vector kernels read their coefficients through pointers into it.
 * ------------------------------------------------------------------------------------------- */
#define PRFDD_KRYLOV_MAXV 32
typedef struct prfdd_krylov_state
{
    double H[(PRFDD_KRYLOV_MAXV + 1) * PRFDD_KRYLOV_MAXV]; /* H[i*MAXV + j] */
    double c[PRFDD_KRYLOV_MAXV];
    double s[PRFDD_KRYLOV_MAXV];
    double gamma[PRFDD_KRYLOV_MAXV + 1];
    double hcol[PRFDD_KRYLOV_MAXV + 1]; /* raw Gram-Schmidt dots of the current column */
    double y[PRFDD_KRYLOV_MAXV];        /* back-substituted coefficients (0 beyond the last column used) */
    double red[8];                      /* scratch for reductions: red[0] norm^2, red[0..1] gamma/theta ... */
    double r0_norm;
    double inv_gamma0;                  /* 1 / gamma[0] */
    double inv_alpha;                   /* 1 / alpha_j */
    double alpha_cg;                    /* inner FCG: alpha_k (0 once stopped) */
    double beta_cg;
    double gamma_cg;
    double r_norm;
    double one;                         /* constant 1.0 (denominator for *_dev kernels) */
    int stopped;                        /* set when a break condition of the reference fired */
    int cycle_active;
    int j_last;                         /* last Arnoldi column whose Z is used */
    int iterations;                     /* accumulated like Subdomain::num_iterations */
} prfdd_krylov_state;

int prfdd_krylov_reset(prfdd_krylov_state *st, prfdd_stream_t stream);           /* new solve: stopped = 0 */
/* outer flexible CG as one CUDA graph (loop = conditional WHILE node): the loop test of domain.tpp:683-695 on the device.
 * state[0] = norms recorded, state[1] = completed search-direction updates; hist[k] = sqrt(*sum) is appended; with has_handle the
 * graph condition `cond_handle` (cudaGraphConditionalHandle) is set to "continue" exactly as the host loop decides */
int prfdd_fcg_outer_reset(int *state, prfdd_stream_t stream);
int prfdd_fcg_outer_check(const double *sum, double *hist, int *state, double tolerance, int use_relative, int max_iterations,
                          unsigned long long cond_handle, int has_handle, prfdd_stream_t stream);
int prfdd_fcg_outer_count(int *state, prfdd_stream_t stream);
/* gamma[0] = sqrt(red[0]); first cycle: r0_norm = gamma[0]; inv_gamma0 = 1/gamma[0]  (tpp:4343-4365) */
int prfdd_gmres_begin_cycle(prfdd_krylov_state *st, int first_cycle, prfdd_stream_t stream);
/* column j: H[0..j][j] = hcol, Givens, alpha = sqrt(red[0]), stop tests of tpp:4415-4453 */
int prfdd_gmres_column(prfdd_krylov_state *st, int j, int iter, int max_iterations, double tolerance,
                       int use_relative, prfdd_stream_t stream);
/* back-substitution (tpp:4460-4470) into y[]; y = 0 for a cycle that started after convergence */
int prfdd_gmres_end_cycle(prfdd_krylov_state *st, int num_vectors, prfdd_stream_t stream);
/* inner FCG scalars: alpha = red[0]/red[1] (gamma/theta) unless stopped                 (tpp:4215) */
int prfdd_fcg_alpha(prfdd_krylov_state *st, prfdd_stream_t stream);
/* r_norm = sqrt(red[2]); stop tests of tpp:4231-4240; iterations++ */
int prfdd_fcg_check(prfdd_krylov_state *st, int iter, int max_iterations, double tolerance, int use_relative,
                    prfdd_stream_t stream);
/* beta = red[3]/gamma (tpp:4256) unless stopped (then beta = 0 and p, r are left untouched by callers) */
int prfdd_fcg_beta(prfdd_krylov_state *st, prfdd_stream_t stream);
/* y += a*x with a read from device memory; nothing happens when *a == 0 exactly */
int prfdd_axpy_dev(double *y, const double *a, double sign, const double *x, int n, prfdd_stream_t stream);
/* p = z + beta p ; r = r1, skipped entirely when *skip_flag != 0                       subdomain.okl:259-266 */
int prfdd_residual_and_search_update_gated(double *p_k, double *r_k, const double *z_k, const double *r_kp1,
                                           const double *beta, const int *skip_flag, int n, prfdd_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * special functions (special_functions.hpp:10-12; host, C++ restatement of Nek5000 speclib)
 * ------------------------------------------------------------------------------------------- */
void prfdd_zwgll(double *z, double *w, int np);
/* D row-major: D[i*n+j] = dl_j/dxi(xi_i) -- what the reference reads out of dgll_ (domain.tpp:312-314) */
void prfdd_dgll(double *D, const double *z, int np);
/* value at x of the j-th (0-based) Lagrange interpolant through the np GLL nodes z */
double prfdd_hgll(int j, double x, const double *z, int np);
/* glibc rand() stream for seed `seed` (TYPE_3 additive feedback), n draws of rand()/RAND_MAX:
 * the reference's function_id 4, domain.tpp:549-550 */
void prfdd_glibc_rand_fill(double *out, long long n, unsigned int seed);

/* ---------------------------------------------------------------------------------------------
 * mesh files (reference on-disk format, domain.tpp:43-224)
 * ------------------------------------------------------------------------------------------- */
/* writes <dir>/lx1_<N+1>/... for every rank of a box mesh [0,1]^dim with nel[d] elements per side,
 * block-partitioned over num_procs ranks (power of two); eps = smooth deformation amplitude */
int prfdd_mesh_generate_box(const char *directory, int dim, const int nel[3], int poly_degree, int num_procs,
                            double eps);
/* the same, writing only the files of rank `only_rank` (every rank of a multi-GPU job writes its own part; -1 = all) */
int prfdd_mesh_generate_box_rank(const char *directory, int dim, const int nel[3], int poly_degree, int num_procs,
                                 double eps, int only_rank);

/* ---------------------------------------------------------------------------------------------
 * halo lists for the process-boundary sum (host logic of Domain::setup_halo; stands where gslib_gs_setup stands,
 * domain.tpp:283-284).  Input: every rank's boundary node ids (rank p: ids[offsets[p] .. offsets[p+1])).
 * Output for rank `proc_id`: its peers in ascending rank order, and for every peer the LOCAL indices (positions in this
 * rank's own id list) of the shared nodes sorted by global id -- both sides of a pair therefore use the same order.
 * peers/peer_count/peer_offset need num_procs entries, idx needs as many entries as this rank has ids times num_procs
 * in the worst case; returns the number of peers (<0 on error), *total = entries written to idx.
 * ------------------------------------------------------------------------------------------- */
int prfdd_halo_build_lists(int proc_id, int num_procs, const long long *ids, const long long *offsets, int *peers, int *peer_count,
                           int *peer_offset, int *idx, long long idx_capacity, long long *total);

/* ---------------------------------------------------------------------------------------------
 * AMG setup on the HOST (no device needed): the library's deterministic stand-in for the two
 * HYPRE_BoomerAMGSetup calls of the reference (subdomain.tpp:1851-1858, 3480-3489).  Exposed so the
 * hierarchy (C/F splittings, interpolation, Galerkin operators, Chebyshev data) can be inspected and
 * compared with the oracle without a GPU.
 * ------------------------------------------------------------------------------------------- */
typedef struct prfdd_amg_host prfdd_amg_host;
int prfdd_amg_host_setup(prfdd_amg_host **h, int n, const int *ptr, const int *col, const double *val, int cheby_order,
                         int max_coarse_size);
/* coarsening: 0 = PMIS with hashed measures, 1 = HMIS (on one process: the first Ruge-Stueben pass, first-in first-out among
 * equal measures) -- what coarsen type 10 of the reference selects (subdomain.tpp:1853); -1 = the library default
 * (PRFDD_AMG_COARSENING=pmis|hmis, else the value DESIGN.md names) */
int prfdd_amg_host_setup_ex(prfdd_amg_host **h, int n, const int *ptr, const int *col, const double *val, int cheby_order,
                            int max_coarse_size, int coarsening);
int prfdd_amg_host_destroy(prfdd_amg_host *h);
int prfdd_amg_host_num_levels(const prfdd_amg_host *h);
/* sizes[0..3] = rows of A_l, nnz of A_l, columns of P_l (0 on the last level), nnz of P_l */
int prfdd_amg_host_level_sizes(const prfdd_amg_host *h, int level, int sizes[4]);
/* which: 0 = A_l, 1 = P_l.  Copies ptr (rows+1), col, val */
int prfdd_amg_host_get_matrix(const prfdd_amg_host *h, int level, int which, int *ptr, int *col, double *val);
/* cf (rows, +1 C / -1 F; empty on the last level), ds (rows), coefs (cheby_order), eigs (max, min) */
int prfdd_amg_host_get_vectors(const prfdd_amg_host *h, int level, signed char *cf, double *ds, double *coefs, double eigs[2]);
/* dense inverse of the coarsest operator, row-major rows x rows */
int prfdd_amg_host_get_coarse_inverse(const prfdd_amg_host *h, double *Ainv);

/* ---------------------------------------------------------------------------------------------
 * solver objects: the reference's run_simulation() sequence (poisson.cpp:150-251) behind handles
 * ------------------------------------------------------------------------------------------- */
typedef struct prfdd_solver prfdd_solver;

typedef struct prfdd_options
{
    int poly_degree;          /* argv[2] */
    int poly_reduction;       /* argv[3] */
    int subdomain_overlap;    /* argv[4] */
    int superdomain_overlap;  /* argv[5] */
    int use_preconditioner;   /* Domain::use_preconditioner, domain.hpp:117 */
    int preconditioner_type;  /* 0 = inner FCG, 1 = inner GMRES (domain.hpp:116) */
    int inner_num_vectors;    /* Subdomain::num_vectors = 4 (subdomain.hpp:229) */
    int inner_max_iterations; /* Subdomain::max_iterations = 4 (subdomain.hpp:230) */
    int num_vcycles;          /* 1 (subdomain.hpp:236) */
    int cheby_order;          /* 2 (subdomain.hpp:237) */
    int use_cuda_graph;       /* AMG/config.hpp:6 USE_CUDA_GRAPH; here the whole preconditioner */
    int proc_id, num_procs;   /* config.hpp:49-50 globals */
    const void *nccl_unique_id; /* 128-byte ncclUniqueId shared by all ranks; NULL when num_procs == 1 */
    double outer_tolerance;   /* Domain::tolerance = 1e-7 (domain.hpp:118) */
    double inner_tolerance;   /* Subdomain::tolerance = 1e-12 (subdomain.hpp:232) */
    int outer_max_iterations; /* 500 (domain.hpp:114) */
    int outer_num_vectors;    /* 20 (domain.hpp:113) */
    int verbose;              /* print the reference's "Iter ..." lines on rank 0 */
    int amg_coarsening;       /* -1 library default, 0 PMIS, 1 HMIS (HYPRE coarsen type 10, subdomain.tpp:1853) */
    int amg_precision;        /* 0 FP64 (`Float double`, the reference's setting, AMG/config.hpp:4), 1 FP32 V-cycle (`Float float`) */
    int device_outer_loop;    /* 1: from the second solve on given buffers the outer flexible CG is ONE CUDA graph whose loop is a
                               * conditional WHILE node with the convergence test on the device -- no host round trip per iteration
                               * (domain.tpp:611-725 reads three reductions per iteration on the host); one rank only.
                               * 0 (default): host-driven loop, one 8-byte read per iteration.  Off by default because it measures
                               * the same (12.80 vs 12.84 ms on c2) and Nsight Compute cannot profile kernel nodes of graphs that
                               * contain conditional nodes */
} prfdd_options;

void prfdd_options_default(prfdd_options *opt);
/* Domain ladder + Subdomain construction (poisson.cpp:172-206) on `stream` of the current device */
int prfdd_solver_create(prfdd_solver **s, const char *directory, const prfdd_options *opt, prfdd_stream_t stream);
int prfdd_solver_destroy(prfdd_solver *s);
/* u* by function_id (domain.tpp:527-580) and f = A_L u* (poisson.cpp:211-219) */
int prfdd_solver_setup_problem(prfdd_solver *s, int function_id);
/* solver_id 0 = flexible CG (domain.tpp:611-725), 1 = flexible GMRES (727-914).  Solves A u = f for the
 * device-resident f; returns iteration count and the residual history (up to history_cap entries). */
int prfdd_solver_solve(prfdd_solver *s, int solver_id, int *num_iterations, double *history, int history_cap,
                       int *history_len);
/* same solve through HOST buffers: f_host (num_local_points) is copied to the device, u is copied back;
 * both copies are inside the call (bench.py "e2e") */
int prfdd_solver_solve_host(prfdd_solver *s, int solver_id, const double *f_host, double *u_host, int *num_iterations,
                            double *history, int history_cap, int *history_len);
/* queries; `what` is one of the PRFDD_Q_* / PRFDD_A_* ids below */
long long prfdd_solver_query(prfdd_solver *s, int what);
/* copies a named host- or device-side array into `dst` (host); returns element count or <0 */
long long prfdd_solver_get_array(prfdd_solver *s, int what, void *dst, long long capacity_bytes);
/* one application of a building block on host buffers, for parity tests */
int prfdd_solver_apply(prfdd_solver *s, int what, const double *in_host, double *out_host);
/* live timing of the dominant kernel of the V-cycle for bench.py's roofline: `reps` back-to-back launches of the fused
 * Chebyshev SpMV step, alternating between AMG levels 0 and 1 (so that no launch finds its matrix in L2), bracketed by CUDA
 * events on the solver's stream.  out[0] = average ms per launch, out[1] = average algorithmic bytes per launch
 * (12*nnz + 4*(rows+1) + 5*8*rows), out[2] = rows of level 0, out[3] = nnz of level 0, out[4] = rows of level 1, out[5] = nnz of level 1 */
int prfdd_solver_time_spmv(prfdd_solver *s, int reps, double out[6]);
/* per-launch profile of the AMG V-cycle of this rank (average of `reps` cycles, every launch bracketed by CUDA events on the solver's
 * stream): a text table, one line per launch -- "level what rows nnz microseconds algorithmic_MB GB/s" -- and a total line.
 * Returns 0, or -(needed capacity). */
int prfdd_solver_profile_vcycle(prfdd_solver *s, int reps, char *text, int capacity);
/* field output (Domain::output, domain.tpp:373-524; the reference writes Silo, compiled out by VISUALIZATION 0): the low-order
 * cell mesh (every GLL cell a quad / hexahedron) with node-centred fields as a legacy-VTK unstructured grid.  Host arrays,
 * element-major points, n = points per direction. */
int prfdd_write_vtk(const char *path, int dim, int n, int num_elements, const double *x, const double *y, const double *z, int num_fields,
                    const char *const *field_names, const double *const *fields);
/* the same for elements of different degrees: element e has n_of_element[e] points per side (Subdomain::output, subdomain.tpp:4648-4791) */
int prfdd_write_vtk_mixed(const char *path, int dim, int num_elements, const int *n_of_element, const double *x, const double *y, const double *z,
                          int num_fields, const char *const *field_names, const double *const *fields);
/* poisson.cpp:233-235: u_star, f, u of this rank's elements -> <output_name>_<rank>.vtk */
int prfdd_solver_output(prfdd_solver *s, const char *output_name);
/* Subdomain::output (subdomain.tpp:4648-4791; never called by the reference's driver): this rank's subdomain region -- own elements
 * at degree N, rings at the ladder degrees, extended N = 1 elements -- as low-order cells with the node fields "degree" (the
 * element's polynomial degree), "f" (the composite right-hand side of the last preconditioner application) and "u" (its
 * solution) -> <output_name>_<rank>.vtk */
int prfdd_solver_output_subdomain(prfdd_solver *s, const char *output_name);
/* timer report (Timer keys of timer.tpp / poisson.cpp:253-401): seconds for `key`, <0 if unknown */
double prfdd_solver_timer_total(prfdd_solver *s, const char *key);

enum
{
    PRFDD_Q_DIM = 1,
    PRFDD_Q_NUM_LOCAL_ELEMENTS,
    PRFDD_Q_NUM_LOCAL_POINTS,
    PRFDD_Q_NUM_LOCAL_NODES,
    PRFDD_Q_NUM_BDARY_NODES,
    PRFDD_Q_NUM_TOTAL_ELEMENTS,
    PRFDD_Q_NUM_GLOBAL_NODES,
    PRFDD_Q_SUB_NUM_POINTS,
    PRFDD_Q_SUB_NUM_DOFS,
    PRFDD_Q_SUB_NUM_EXTENDED_DOFS,
    PRFDD_Q_SUP_NUM_DOFS,
    PRFDD_Q_SUP_NUM_EXTENDED_DOFS,
    PRFDD_Q_NUM_VALUES,
    PRFDD_Q_NUM_DOFS,
    PRFDD_Q_AMG_NUM_LEVELS,
    PRFDD_Q_INNER_ITERATIONS,
    PRFDD_Q_GPU_LAUNCHES_PER_PRECOND
};

enum
{
    PRFDD_A_NODE_OF_POINT = 100, /* int32[P]   Q.col : local node of every point (domain.tpp:249-291) */
    PRFDD_A_BOUNDARY_NODES,      /* int64[nb]  gs ids of the process-boundary nodes (domain.tpp:261) */
    PRFDD_A_ASSEMBLED_WEIGHT,    /* f64[Nn]    1/multiplicity (domain.tpp:296-302) */
    PRFDD_A_D_HAT,               /* f64[n*n] */
    PRFDD_A_U,                   /* f64[P] solution */
    PRFDD_A_U_STAR,              /* f64[P] */
    PRFDD_A_F,                   /* f64[P] */
    PRFDD_A_SUB_Q_PTR,           /* region Q (points x extended dofs) CSR */
    PRFDD_A_SUB_Q_COL,
    PRFDD_A_SUB_Q_VAL,
    PRFDD_A_SUB_ELEMENT_IDS,     /* int32[] global element id of every region element, in region order */
    PRFDD_A_SUB_ELEMENT_DEGREE,  /* int32[] */
    PRFDD_A_SUB_DOF_NUM,         /* int64[num_points] dof_num of every region point (0 = masked / slave) */
    PRFDD_A_AMG_LEVEL_ROWS,      /* int32[levels] */
    PRFDD_A_AMG_LEVEL_NNZ,       /* int32[levels] */
    PRFDD_A_AMG_CHEBY_COEFS,     /* f64[levels*cheby_order] */
    PRFDD_A_A_FEM_PTR,           /* level-0 low-order FEM matrix CSR */
    PRFDD_A_A_FEM_COL,
    PRFDD_A_A_FEM_VAL,
    PRFDD_A_NORM_WEIGHT,
    PRFDD_A_INNER_WEIGHT
};

enum
{
    PRFDD_APPLY_STIFFNESS = 200,      /* Domain::stiffness_matrix        in[P] -> out[P] */
    PRFDD_APPLY_DSSUM,                /* dssum(mask=true, weight=false)  in[P] -> out[P] */
    PRFDD_APPLY_DSSUM_WEIGHTED,       /* dssum(mask=true, weight=true) */
    PRFDD_APPLY_PRECONDITIONER,       /* z = M^-1 r (inner Krylov, before stitching) in[P] -> out[P] */
    PRFDD_APPLY_SUB_STIFFNESS,        /* Subdomain::stiffness_matrix     in[num_values] -> out[num_values] */
    PRFDD_APPLY_LOW_ORDER,            /* Subdomain::low_order_preconditioner in/out[num_values] */
    PRFDD_APPLY_VCYCLE,               /* one V-cycle on level-0 dofs     in[num_dofs] -> out[num_dofs] */
    PRFDD_APPLY_TREE                  /* Subdomain::tree_operator        in[P] -> out[num_values] */
};

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PRFDD_B200_H */
