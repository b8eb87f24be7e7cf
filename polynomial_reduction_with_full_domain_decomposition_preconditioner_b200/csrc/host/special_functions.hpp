// special_functions.hpp -- the three speclib entry points the reference calls
// (/root/reference/special_functions.hpp:10-12), same names and calling convention (Fortran:
// everything by reference), implemented in C++ in special_functions.cpp.
#pragma once
extern "C"
{
    void zwgll_(double *z, double *w, const int *np);
    void dgll_(double *d, double *dt, double *z, const int *nz, const int *lzd);
    double hgll_(const int *i, double *z, double *zgll, const int *nz);
}
