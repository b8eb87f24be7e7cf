// math.hpp -- Math<DType>: the reference's four vector kernels behind the same method names
// (/root/reference/math.hpp:14-34, math.tpp:45-67, math.okl:5-35).  matrix_matrix_multiply
// (math.tpp:69-92) is dead code upstream (never called, and mis-indexed) and is not carried over.
#pragma once
#include "config.hpp"
#include "../../../include/prfdd_b200.h"

template <typename DType>
class Math
{
  public:
    Math() {}
    ~Math() {}

    void set_to_value(const dev::memory &u, DType alpha, int n, int offset = 0)
    {
        dev::check_rc(prfdd_set_to_value(u.as<double>(), alpha, n, offset, prfdd_host::device.stream), "Math::set_to_value");
    }
    void invert_vector_elements(const dev::memory &u, int n)
    {
        dev::check_rc(prfdd_invert_vector_elements(u.as<double>(), n, prfdd_host::device.stream), "Math::invert_vector_elements");
    }
    void vector_vector_addition(const dev::memory &uv, const DType alpha, const dev::memory &u, const DType beta, const dev::memory &v, const int n)
    {
        dev::check_rc(prfdd_vector_vector_addition(uv.as<double>(), alpha, u.as<double>(), beta, v.as<double>(), n, prfdd_host::device.stream), "Math::vector_vector_addition");
    }
    void vector_scaling(const dev::memory &au, const DType alpha, const dev::memory &u, const int n)
    {
        dev::check_rc(prfdd_vector_scaling(au.as<double>(), alpha, u.as<double>(), n, prfdd_host::device.stream), "Math::vector_scaling");
    }
};
