// subdomain.hpp -- placeholder until the PR-FDD preconditioner lands (next commit)
#pragma once
#include "config.hpp"
#include "domain.hpp"
template <typename DType>
class Subdomain
{
  public:
    Subdomain() {}
    template <typename Map>
    Subdomain(Map &, int, int, int, int, const prfdd_options &) { throw std::runtime_error("Subdomain: not built yet"); }
    void flexible_conjugate_gradient(dev::memory &, dev::memory &) {}
    void generalized_minimum_residual(dev::memory &, dev::memory &) {}
    long long query(int) { return -1; }
    long long get_array(int, void *, long long) { return -1; }
    int apply(int, const double *, double *) { return -1; }
};
