// special_functions.cpp -- C++ restatement of the Nek5000 speclib routines reached by zwgll_, dgll_
// and hgll_ (/root/reference/special_functions.f: ZWGLL 108-123, ZWGLJD 233-273, ZWGJD 153-201,
// JACG 414-459, JACOBF 461-500, ENDW1/2 276-352, GAMMAF 354-377, PNORMJ 379-402, DGLL 781-814,
// HGLL 816-835, PNLEG 856-884, PNDLEG 886-913).  The reference builds that file with
// -fdefault-real-8 (Makefile:53): all arithmetic is double, every literal is a double literal.
// The arithmetic is kept operation for operation so that the GLL nodes, D and J matrices carry the
// same bits the reference's would.
#include "special_functions.hpp"
#include "../../../include/prfdd_b200.h"
#include <cmath>
#include <vector>

namespace
{
// Jacobi polynomial P_n^(a,b) with derivative and the two previous degrees (JACOBF)
struct Jacobi
{
    double p = 1.0, pd = 0.0, pm1 = 0.0, pdm1 = 0.0, pm2 = 0.0, pdm2 = 0.0;
    Jacobi(int n, double a, double b, double x)
    {
        const double apb = a + b;
        if (n == 0) return;
        double pl = p, pdl = pd, ps = 0.0, pds = 0.0;
        p = (a - b + (apb + 2.) * x) / 2.;
        pd = (apb + 2.) / 2.;
        if (n == 1) return;
        for (int k = 2; k <= n; k++)
        {
            const double dk = (double)k;
            const double a1 = 2. * dk * (dk + apb) * (2. * dk + apb - 2.);
            const double a2 = (2. * dk + apb - 1.) * (a * a - b * b);
            const double b3 = (2. * dk + apb - 2.);
            const double a3 = b3 * (b3 + 1.) * (b3 + 2.);
            const double a4 = 2. * (dk + a - 1.) * (dk + b - 1.) * (2. * dk + apb);
            const double pn = ((a2 + a3 * x) * p - a4 * pl) / a1;
            const double pdn = ((a2 + a3 * x) * pd - a4 * pdl + a3 * p) / a1;
            ps = pl; pds = pdl;
            pl = p; p = pn;
            pdl = pd; pd = pdn;
        }
        pm1 = pl; pdm1 = pdl; pm2 = ps; pdm2 = pds;
    }
};

double gammaf(double x)
{
    const double pi = 4.0 * std::atan(1.0);
    double g = 1.0;
    if (x == -0.5) g = -2.0 * std::sqrt(pi);
    if (x == 0.5) g = std::sqrt(pi);
    if (x == 1.0) g = 1.0;
    if (x == 2.0) g = 1.0;
    if (x == 1.5) g = std::sqrt(pi) / 2.;
    if (x == 2.5) g = 1.5 * std::sqrt(pi) / 2.;
    if (x == 3.5) g = 0.5 * (2.5 * (1.5 * std::sqrt(pi)));
    if (x == 3.) g = 2.;
    if (x == 4.) g = 6.;
    if (x == 5.) g = 24.;
    if (x == 6.) g = 120.;
    return g;
}

double pnormj(int n, double a, double b)
{
    const double dn = (double)n, c = a + b + 1.0;
    double prod;
    if (n <= 1)
    {
        prod = gammaf(dn + a) * gammaf(dn + b);
        prod = prod / (gammaf(dn) * gammaf(dn + a + b));
        return prod * std::pow(2.0, c) / (2.0 * dn + c);
    }
    prod = gammaf(a + 1.0) * gammaf(b + 1.0);
    prod = prod / (2.0 * (1.0 + c) * gammaf(c + 1.0));
    prod = prod * (1.0 + a) * (2.0 + a);
    prod = prod * (1.0 + b) * (2.0 + b);
    for (int i = 3; i <= n; i++)
    {
        const double di = (double)i;
        prod = prod * ((di + a) * (di + b) / (di * (di + a + b)));
    }
    return prod * std::pow(2.0, c) / (2.0 * dn + c);
}

// end weights of the Lobatto rule; which = 1 (left, ENDW1) or 2 (right, ENDW2)
double endw(int which, int n, double a, double b)
{
    const double apb = a + b;
    if (n == 0) return 0.;
    double f1 = (which == 1) ? gammaf(a + 2.) * gammaf(b + 1.) / gammaf(apb + 3.) : gammaf(a + 1.) * gammaf(b + 2.) / gammaf(apb + 3.);
    f1 = f1 * (apb + 2.) * std::pow(2., apb + 2.) / 2.;
    if (n == 1) return f1;
    double fint1 = (which == 1) ? gammaf(a + 2.) * gammaf(b + 1.) / gammaf(apb + 3.) : gammaf(a + 1.) * gammaf(b + 2.) / gammaf(apb + 3.);
    fint1 = fint1 * std::pow(2., apb + 2.);
    double fint2 = gammaf(a + 2.) * gammaf(b + 2.) / gammaf(apb + 4.);
    fint2 = fint2 * std::pow(2., apb + 3.);
    double f2 = (which == 1) ? (-2. * (b + 2.) * fint1 + (apb + 4.) * fint2) * (apb + 3.) / 4.
                             : (2. * (a + 2.) * fint1 - (apb + 4.) * fint2) * (apb + 3.) / 4.;
    if (n == 2) return f2;
    double f3 = 0.;
    for (int i = 3; i <= n; i++)
    {
        const double di = (double)(i - 1);
        const double abn = a + b + di, abnn = abn + di;
        const double a1 = -(2. * (di + a) * (di + b)) / (abn * abnn * (abnn + 1.));
        const double a2 = (2. * (a - b)) / (abnn * (abnn + 2.));
        const double a3 = (2. * (abn + 1.)) / ((abnn + 2.) * (abnn + 1.));
        f3 = -(a2 * f2 + a1 * f1) / a3;
        f1 = f2;
        f2 = f3;
    }
    return f3;
}

// zeros of P_np^(a,b): Newton with deflation, at most 10 steps, eps 1e-12, then selection sort (JACG)
void jacg(double *x, int np, double a, double b)
{
    const int n = np - 1;
    const double one = 1.;
    const double dth = 4. * std::atan(one) / (2. * ((double)n) + 2.);
    double xc = 0., xlast = 0.;
    for (int j = 1; j <= np; j++)
    {
        if (j == 1)
            xc = std::cos((2. * (((double)j) - 1.) + 1.) * dth);
        else
            xc = (std::cos((2. * (((double)j) - 1.) + 1.) * dth) + xlast) / 2.;
        for (int k = 1; k <= 10; k++)
        {
            Jacobi q(np, a, b, xc);
            double recsum = 0.;
            for (int i = 1; i <= j - 1; i++) recsum = recsum + 1. / (xc - x[np - i]);
            const double delx = -q.p / (q.pd - recsum * q.p);
            xc = xc + delx;
            if (std::fabs(delx) < 1.0e-12) break;
        }
        x[np - j] = xc;
        xlast = xc;
    }
    for (int i = 0; i < np; i++)
    {
        double xmin = 2.;
        int jmin = i;
        for (int j = i; j < np; j++)
            if (x[j] < xmin) { xmin = x[j]; jmin = j; }
        if (jmin != i) { const double t = x[i]; x[i] = x[jmin]; x[jmin] = t; }
    }
}

void zwgjd(double *z, double *w, int np, double a, double b)
{
    const int n = np - 1;
    const double apb = a + b;
    if (np == 1)
    {
        z[0] = (b - a) / (apb + 2.);
        w[0] = gammaf(a + 1.) * gammaf(b + 1.) / gammaf(apb + 2.) * std::pow(2., apb + 1.);
        return;
    }
    jacg(z, np, a, b);
    const int np1 = n + 1, np2 = n + 2;
    const double dnp1 = (double)np1, dnp2 = (double)np2;
    const double fac1 = dnp1 + a + b + 1., fac2 = fac1 + dnp1, fac3 = fac2 + 1.;
    const double fnorm = pnormj(np1, a, b);
    const double rcoef = (fnorm * fac2 * fac3) / (2. * fac1 * dnp2);
    for (int i = 0; i < np; i++)
    {
        Jacobi q(np2, a, b, z[i]);
        w[i] = -rcoef / (q.p * q.pdm1);
    }
}

// Legendre P_n at *z; snaps |*z| < 1e-25 to exactly 0 in the caller's storage (PNLEG, f:868)
double pnleg(double *z, int n)
{
    if (std::fabs(*z) < 1.0e-25) *z = 0.0;
    double p1 = 1.;
    if (n == 0) return p1;
    double p2 = *z, p3 = p2;
    for (int k = 1; k <= n - 1; k++)
    {
        const double fk = (double)k;
        p3 = ((2. * fk + 1.) * (*z) * p2 - fk * p1) / (fk + 1.);
        p1 = p2;
        p2 = p3;
    }
    return p3;
}

double pndleg(double z, int n)
{
    double p1 = 1., p2 = z, p1d = 0., p2d = 1., p3d = 1.;
    for (int k = 1; k <= n - 1; k++)
    {
        const double fk = (double)k;
        const double p3 = ((2. * fk + 1.) * z * p2 - fk * p1) / (fk + 1.);
        p3d = ((2. * fk + 1.) * p2 + (2. * fk + 1.) * z * p2d - fk * p1d) / (fk + 1.);
        p1 = p2; p2 = p3; p1d = p2d; p2d = p3d;
    }
    return (n == 0) ? 0. : p3d;
}
} // namespace

extern "C" {

void zwgll_(double *z, double *w, const int *np_)
{
    const int np = *np_, n = np - 1, nm1 = n - 1;
    const double a = 0., b = 0.;
    if (nm1 > 0) zwgjd(z + 1, w + 1, nm1, a + 1., b + 1.);
    z[0] = -1.;
    z[np - 1] = 1.;
    for (int i = 1; i < np - 1; i++) w[i] = w[i] / (1. - z[i] * z[i]);
    {
        Jacobi q(n, a, b, z[0]);
        w[0] = endw(1, n, a, b) / (2. * q.pd);
    }
    {
        Jacobi q(n, a, b, z[np - 1]);
        w[np - 1] = endw(2, n, a, b) / (2. * q.pd);
    }
}

void dgll_(double *d, double *dt, double *z, const int *nz_, const int *lzd_)
{
    const int nz = *nz_, lzd = *lzd_, n = nz - 1;
    if (nz == 1) { d[0] = 0.; return; }
    const double fn = (double)n, d0 = fn * (fn + 1.) / 4.;
    for (int i = 1; i <= nz; i++)
        for (int j = 1; j <= nz; j++)
        {
            double v = 0.;
            if (i != j) v = pnleg(&z[i - 1], n) / (pnleg(&z[j - 1], n) * (z[i - 1] - z[j - 1]));
            if (i == j && i == 1) v = -d0;
            if (i == j && i == nz) v = d0;
            d[(i - 1) + (j - 1) * lzd] = v;   // Fortran D(I,J)
            dt[(j - 1) + (i - 1) * lzd] = v;  // Fortran DT(J,I)
        }
}

double hgll_(const int *ii, double *z, double *zgll, const int *nz_)
{
    const int i = *ii, nz = *nz_;
    const double dz = *z - zgll[i - 1];
    if (std::fabs(dz) < 1.e-5) return 1.;
    const int n = nz - 1;
    const double alfan = ((double)n) * (((double)n) + 1.);
    return -(1. - (*z) * (*z)) * pndleg(*z, n) / (alfan * pnleg(&zgll[i - 1], n) * (*z - zgll[i - 1]));
}

// ---- C ABI conveniences (include/prfdd_b200.h) ----
void prfdd_zwgll(double *z, double *w, int np) { zwgll_(z, w, &np); }

void prfdd_dgll(double *D, const double *z, int np)
{
    std::vector<double> zz(z, z + np), Dt(np * np);
    // the reference passes (Dt_gll, D_gll): its row-major read of the second array is D[i][j] (domain.tpp:312-314)
    dgll_(Dt.data(), D, zz.data(), &np, &np);
}

double prfdd_hgll(int j, double x, const double *z, int np)
{
    std::vector<double> zz(z, z + np);
    int jj = j + 1;
    return hgll_(&jj, &x, zz.data(), &np);
}

// glibc random_r TYPE_3 (degree 31, separation 3): the stream rand() yields after srand(seed)
void prfdd_glibc_rand_fill(double *out, long long n, unsigned int seed)
{
    std::vector<int32_t> r(34 + 310);
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++)
    {
        // 16807 * r[i-1] % 2147483647 without overflow (Schrage), as glibc's srandom_r does
        int64_t hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        int64_t word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = (int32_t)word;
    }
    // ring buffer of 31 words, front pointer 3 ahead of rear
    uint32_t s[31];
    for (int i = 0; i < 31; i++) s[i] = (uint32_t)r[i];
    int f = 3, rr = 0;
    auto next = [&]() -> uint32_t {
        s[f] += s[rr];
        uint32_t result = s[f] >> 1;
        f = (f + 1) % 31;
        rr = (rr + 1) % 31;
        return result;
    };
    for (int i = 0; i < 310; i++) next();
    for (long long k = 0; k < n; k++) out[k] = (double)(next()) / (double)(2147483647);
}

} // extern "C"
