/*
 * prfdd_compat.h -- the reference's own names over libprfdd_b200.so, for a maintainer who swaps the library in.
 *
 * 1. The reference's C ABI (AMG/kernels.cu:18-94): libprfdd_compat.so exports these five symbols with the reference's exact
 *    names and argument lists; its host code (subdomain.tpp:17, 42-43, 70; AMG/vector.cpp:71) links unchanged.
 * 2. The OKL kernels whose meaning carries over one to one (math.okl, csr_matrix.okl, the element-wise kernels of domain.okl /
 *    subdomain.okl): C++ inline functions in namespace prfdd_okl with the kernel's name and argument order; an OCCA call
 *    `kernel(args...)` becomes `prfdd_okl::kernel(args..., stream)` with `occa::memory::ptr()` pointers.
 * What has NO one-to-one form, and why (INTEGRATION.md shows the call-site edits):
 *    stiffness_matrix_1 + stiffness_matrix_2 (domain.okl:5-98, subdomain.okl:4-101) are one fused launch here
 *      (prfdd_stiffness_matrix_hd / prfdd_stiffness_matrix_region_hd): the GDu temporaries do not exist;
 *    restriction_1/2/3 (subdomain.okl:284-366) are one launch (prfdd_restriction);
 *    the reduction kernels write ONE finished sum per quantity into device memory instead of `num_blocks` block partials that
 *      the host adds (domain.tpp:916-996): the wrappers below keep the kernel's argument list and put the finished sums in
 *      block[0] (and block[num_blocks] for the second sum of the two-sum kernels), so a host loop over `num_blocks` partials
 *      must be replaced by reading those entries.
 */
#ifndef PRFDD_COMPAT_H
#define PRFDD_COMPAT_H

#include <cuda_runtime.h>
#include "prfdd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
/* AMG/kernels.cu:18, 36, 54, 71, 89 (Float = double, AMG/config.hpp:4) -- exported by libprfdd_compat.so */
void vector_set_to_value(double *data, const double value, const int size, cudaStream_t stream);
void main_scaled_residual(double *Sr, double *w, const double *f_m_Au, const double *S, const double alpha, const int size, cudaStream_t stream);
void main_polynomial_evaluation(double *w, double *v, const double *r, const double *D_val, const double alpha, const int size, cudaStream_t stream);
void main_update_field(double *u, const double *w, const double *D_val, const int size, cudaStream_t stream);
void vector_multiplication(double *uv, const double *u, const double *v, const int size, cudaStream_t stream);
#ifdef __cplusplus
}

namespace prfdd_okl
{
typedef double DType;
typedef double EType;
/* math.okl:5-35 */
inline int set_to_value(DType *u, DType alpha, int n, int offset, cudaStream_t s) { return prfdd_set_to_value(u, alpha, n, offset, s); }
inline int invert_vector_elements(DType *u, int n, cudaStream_t s) { return prfdd_invert_vector_elements(u, n, s); }
inline int vector_vector_addition(DType *uv, DType alpha, const DType *u, DType beta, const DType *v, int n, cudaStream_t s) { return prfdd_vector_vector_addition(uv, alpha, u, beta, v, n, s); }
inline int vector_scaling(DType *au, DType alpha, const DType *u, int n, cudaStream_t s) { return prfdd_vector_scaling(au, alpha, u, n, s); }
/* csr_matrix.okl:5-48 (row_end inclusive, as in the reference) */
inline int multiply(DType *Au, const int *ptr, const int *col, const DType *val, const DType *u, int n, cudaStream_t s) { return prfdd_csr_multiply(Au, ptr, col, val, u, n, 0, s); }
inline int multiply_range(DType *Au, const int *ptr, const int *col, const DType *val, const DType *u, int row_start, int row_end, cudaStream_t s) { return prfdd_csr_multiply_range(Au, ptr, col, val, u, row_start, row_end, 0, s); }
inline int multiply_weight(DType *Au, const int *ptr, const int *col, const DType *val, const DType *u, const DType *weight, int n, cudaStream_t s) { return prfdd_csr_multiply_weight(Au, ptr, col, val, u, weight, n, 0, s); }
/* domain.okl:100-107, 186-193, 226-233 and their subdomain.okl twins */
inline int initialize_arrays(DType *u_k, DType *r_k, const DType *f, int n, cudaStream_t s) { return prfdd_initialize_arrays(u_k, r_k, f, n, s); }
inline int solution_and_residual_update(DType *u_k, DType *r_kp1, const DType *r_k, const DType *p_k, const DType *q_k, DType alpha_k, int n, cudaStream_t s) { return prfdd_solution_and_residual_update(u_k, r_kp1, r_k, p_k, q_k, alpha_k, n, s); }
inline int residual_and_search_update(DType *p_k, DType *r_k, const DType *z_k, const DType *r_kp1, DType beta_k, int n, cudaStream_t s) { return prfdd_residual_and_search_update(p_k, r_k, z_k, r_kp1, beta_k, n, s); }
/* subdomain.okl:268-282 */
inline int copy_from_domain_data(DType *u, const EType *v, int n, cudaStream_t s) { return prfdd_copy_from_domain_data(u, v, n, s); }
inline int copy_to_domain_data(EType *u, const DType *v, int n, cudaStream_t s) { return prfdd_copy_to_domain_data(u, v, n, s); }
/* reductions (domain.okl:109-264): finished sums in block[0] (and block[num_blocks]); ws from prfdd_reduce_ws_create */
inline int residual_norm(prfdd_reduce_ws *ws, DType *block, const DType *r_k, const DType *QQt_r_k, const DType *mask, int n, int /*num_blocks*/, cudaStream_t s) { return prfdd_residual_norm(ws, block, r_k, QQt_r_k, mask, n, s); }
inline int inner_product(prfdd_reduce_ws *ws, DType *block, const DType *u_k, const DType *v_k, const DType *mask, int n, int /*num_blocks*/, cudaStream_t s) { return prfdd_inner_product(ws, block, u_k, v_k, mask, n, s); }
inline int inner_product_flexible(prfdd_reduce_ws *ws, DType *block, const DType *r_k, const DType *r_kp1, const DType *z_k, int n, int /*num_blocks*/, cudaStream_t s) { return prfdd_inner_product_flexible(ws, block, r_k, r_kp1, z_k, n, s); }
} // namespace prfdd_okl
#endif

#endif /* PRFDD_COMPAT_H */
