// compat.cpp -- libprfdd_compat.so: the reference's OWN C ABI, symbol for symbol.
//
// The reference declares five `extern "C" void` launchers (AMG/kernels.cu:18, 36, 54, 71, 89; used from subdomain.tpp:17, 42-43, 70
// and AMG/vector.cpp:71) and links them from its CUDA object.  This small library exports exactly those names and argument lists
// and forwards to libprfdd_b200.so, so the reference's host code links against it unchanged (drop AMG/kernels.cu from the link
// line, add -lprfdd_compat -lprfdd_b200).  They live in their own library because the names are generic.
#include <cuda_runtime.h>
#include "../../include/prfdd_b200.h"

#if defined(__GNUC__)
#define PRFDD_EXPORT __attribute__((visibility("default")))
#else
#define PRFDD_EXPORT
#endif

typedef double Float; // AMG/config.hpp:4

extern "C" {

PRFDD_EXPORT void vector_set_to_value(Float *data, const Float value, const int size, cudaStream_t stream)
{
    prfdd_vector_set_to_value(data, value, size, (prfdd_stream_t)stream);
}

PRFDD_EXPORT void main_scaled_residual(Float *Sr, Float *w, const Float *f_m_Au, const Float *S, const Float alpha, const int size, cudaStream_t stream)
{
    prfdd_main_scaled_residual(Sr, w, f_m_Au, S, alpha, size, (prfdd_stream_t)stream);
}

PRFDD_EXPORT void main_polynomial_evaluation(Float *w, Float *v, const Float *r, const Float *D_val, const Float alpha, const int size, cudaStream_t stream)
{
    prfdd_main_polynomial_evaluation(w, v, r, D_val, alpha, size, (prfdd_stream_t)stream);
}

PRFDD_EXPORT void main_update_field(Float *u, const Float *w, const Float *D_val, const int size, cudaStream_t stream)
{
    prfdd_main_update_field(u, w, D_val, size, (prfdd_stream_t)stream);
}

PRFDD_EXPORT void vector_multiplication(Float *uv, const Float *u, const Float *v, const int size, cudaStream_t stream)
{
    prfdd_vector_multiplication(uv, u, v, size, (prfdd_stream_t)stream);
}

} // extern "C"
