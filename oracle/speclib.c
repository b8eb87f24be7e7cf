/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * CPU restatement in plain C of the three Nek5000 speclib entry points the
 * reference calls -- zwgll_, dgll_, hgll_ -- and of the routines they reach.
 * The reference compiles special_functions.f with `-fdefault-real-8`
 * (Makefile:53), so every REAL and every real literal below is a double.
 *
 * Follows /root/reference/special_functions.f:
 *   ZWGLL 108-123, ZWGJD 153-201, ZWGLJD 233-273, ENDW1 276-313, ENDW2 315-352,
 *   GAMMAF 354-377, PNORMJ 379-402, JACG 404-459, JACOBF 461-500,
 *   DGLL 781-814, HGLL 816-835, PNLEG 856-884, PNDLEG 886-913.
 * Call conventions restated from domain.tpp:305-316 and subdomain.tpp:130-184.
 *
 * Parity status: the Fortran cannot be compiled here (no gfortran), so this
 * file is pinned only by mathematical known answers (closed-form GLL nodes for
 * n<=5, numpy Legendre roots, exact differentiation of polynomials): see
 * tests/test_oracle_speclib.py.
 *
 * Fortran passes everything by reference; PNLEG *writes* to its Z argument
 * (special_functions.f:868: values |Z|<1e-25 are snapped to 0), which changes the
 * caller's node array in place.  That side effect is kept: pnleg takes a pointer.
 */
#include <math.h>
#include <stdio.h>

static double gammaf(double x)
{
    const double pi = 4.0 * atan(1.0);
    double g = 1.0;
    if (x == -0.5) g = -2.0 * sqrt(pi);
    if (x == 0.5) g = sqrt(pi);
    if (x == 1.0) g = 1.0;
    if (x == 2.0) g = 1.0;
    if (x == 1.5) g = sqrt(pi) / 2.;
    if (x == 2.5) g = 1.5 * sqrt(pi) / 2.;
    if (x == 3.5) g = 0.5 * (2.5 * (1.5 * sqrt(pi)));
    if (x == 3.) g = 2.;
    if (x == 4.) g = 6.;
    if (x == 5.) g = 24.;
    if (x == 6.) g = 120.;
    return g;
}

static double pnormj(int n, double alpha, double beta)
{
    double dn = (double)n;
    double cnst = alpha + beta + 1.0;
    double prod;
    if (n <= 1)
    {
        prod = gammaf(dn + alpha) * gammaf(dn + beta);
        prod = prod / (gammaf(dn) * gammaf(dn + alpha + beta));
        return prod * pow(2.0, cnst) / (2.0 * dn + cnst);
    }
    prod = gammaf(alpha + 1.0) * gammaf(beta + 1.0);
    prod = prod / (2.0 * (1.0 + cnst) * gammaf(cnst + 1.0));
    prod = prod * (1.0 + alpha) * (2.0 + alpha);
    prod = prod * (1.0 + beta) * (2.0 + beta);
    for (int i = 3; i <= n; i++)
    {
        double dindx = (double)i;
        double frac = (dindx + alpha) * (dindx + beta) / (dindx * (dindx + alpha + beta));
        prod = prod * frac;
    }
    return prod * pow(2.0, cnst) / (2.0 * dn + cnst);
}

/* JACOBF: Jacobi polynomial of degree n and derivative at x (+ the two lower degrees). */
static void jacobf(double *poly, double *pder, double *polym1, double *pderm1, double *polym2, double *pderm2,
                   int n, double alp, double bet, double x)
{
    double apb = alp + bet;
    double polyl, pderl, psave = 0.0, pdsave = 0.0;
    *poly = 1.;
    *pder = 0.;
    if (n == 0) return;
    polyl = *poly;
    pderl = *pder;
    *poly = (alp - bet + (apb + 2.) * x) / 2.;
    *pder = (apb + 2.) / 2.;
    if (n == 1) return;
    for (int k = 2; k <= n; k++)
    {
        double dk = (double)k;
        double a1 = 2. * dk * (dk + apb) * (2. * dk + apb - 2.);
        double a2 = (2. * dk + apb - 1.) * (alp * alp - bet * bet);
        double b3 = (2. * dk + apb - 2.);
        double a3 = b3 * (b3 + 1.) * (b3 + 2.);
        double a4 = 2. * (dk + alp - 1.) * (dk + bet - 1.) * (2. * dk + apb);
        double polyn = ((a2 + a3 * x) * (*poly) - a4 * polyl) / a1;
        double pdern = ((a2 + a3 * x) * (*pder) - a4 * pderl + a3 * (*poly)) / a1;
        psave = polyl;
        pdsave = pderl;
        polyl = *poly;
        *poly = polyn;
        pderl = *pder;
        *pder = pdern;
    }
    *polym1 = polyl;
    *pderm1 = pderl;
    *polym2 = psave;
    *pderm2 = pdsave;
}

/* JACG: the np zeros of the Jacobi polynomial, Newton with deflation, <=10 steps, eps 1e-12. */
static void jacg(double *xjac, int np, double alpha, double beta)
{
    const int kstop = 10;
    const double eps = 1.0e-12;
    int n = np - 1;
    double one = 1.;
    double dth = 4. * atan(one) / (2. * ((double)n) + 2.);
    double x = 0.0, xlast = 0.0;
    double p, pd, pm1, pdm1, pm2, pdm2;
    for (int j = 1; j <= np; j++)
    {
        if (j == 1)
        {
            x = cos((2. * (((double)j) - 1.) + 1.) * dth);
        }
        else
        {
            double x1 = cos((2. * (((double)j) - 1.) + 1.) * dth);
            double x2 = xlast;
            x = (x1 + x2) / 2.;
        }
        for (int k = 1; k <= kstop; k++)
        {
            jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, np, alpha, beta, x);
            double recsum = 0.;
            int jm = j - 1;
            for (int i = 1; i <= jm; i++) recsum = recsum + 1. / (x - xjac[(np - i + 1) - 1]);
            double delx = -p / (pd - recsum * p);
            x = x + delx;
            if (fabs(delx) < eps) break;
        }
        xjac[(np - j + 1) - 1] = x;
        xlast = x;
    }
    for (int i = 1; i <= np; i++)
    {
        double xmin = 2.;
        int jmin = i;
        for (int j = i; j <= np; j++)
        {
            if (xjac[j - 1] < xmin)
            {
                xmin = xjac[j - 1];
                jmin = j;
            }
        }
        if (jmin != i)
        {
            double swap = xjac[i - 1];
            xjac[i - 1] = xjac[jmin - 1];
            xjac[jmin - 1] = swap;
        }
    }
}

/* ZWGJD: Gauss-Jacobi points and weights. */
static void zwgjd(double *z, double *w, int np, double alpha, double beta)
{
    int n = np - 1;
    double one = 1., two = 2.;
    double apb = alpha + beta;
    if (np == 1)
    {
        z[0] = (beta - alpha) / (apb + two);
        w[0] = gammaf(alpha + one) * gammaf(beta + one) / gammaf(apb + two) * pow(two, apb + one);
        return;
    }
    jacg(z, np, alpha, beta);
    int np1 = n + 1;
    int np2 = n + 2;
    double dnp1 = (double)np1;
    double dnp2 = (double)np2;
    double fac1 = dnp1 + alpha + beta + one;
    double fac2 = fac1 + dnp1;
    double fac3 = fac2 + one;
    double fnorm = pnormj(np1, alpha, beta);
    double rcoef = (fnorm * fac2 * fac3) / (two * fac1 * dnp2);
    for (int i = 0; i < np; i++)
    {
        double p, pd, pm1, pdm1, pm2, pdm2;
        jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, np2, alpha, beta, z[i]);
        w[i] = -rcoef / (p * pdm1);
    }
}

static double endw1(int n, double alpha, double beta)
{
    double zero = 0., one = 1., two = 2., three = 3., four = 4.;
    double apb = alpha + beta;
    double f1, f2, f3 = 0.0, fint1, fint2;
    if (n == 0) return zero;
    f1 = gammaf(alpha + two) * gammaf(beta + one) / gammaf(apb + three);
    f1 = f1 * (apb + two) * pow(two, apb + two) / two;
    if (n == 1) return f1;
    fint1 = gammaf(alpha + two) * gammaf(beta + one) / gammaf(apb + three);
    fint1 = fint1 * pow(two, apb + two);
    fint2 = gammaf(alpha + two) * gammaf(beta + two) / gammaf(apb + four);
    fint2 = fint2 * pow(two, apb + three);
    f2 = (-two * (beta + two) * fint1 + (apb + four) * fint2) * (apb + three) / four;
    if (n == 2) return f2;
    for (int i = 3; i <= n; i++)
    {
        double di = (double)(i - 1);
        double abn = alpha + beta + di;
        double abnn = abn + di;
        double a1 = -(two * (di + alpha) * (di + beta)) / (abn * abnn * (abnn + one));
        double a2 = (two * (alpha - beta)) / (abnn * (abnn + two));
        double a3 = (two * (abn + one)) / ((abnn + two) * (abnn + one));
        f3 = -(a2 * f2 + a1 * f1) / a3;
        f1 = f2;
        f2 = f3;
    }
    return f3;
}

static double endw2(int n, double alpha, double beta)
{
    double zero = 0., one = 1., two = 2., three = 3., four = 4.;
    double apb = alpha + beta;
    double f1, f2, f3 = 0.0, fint1, fint2;
    if (n == 0) return zero;
    f1 = gammaf(alpha + one) * gammaf(beta + two) / gammaf(apb + three);
    f1 = f1 * (apb + two) * pow(two, apb + two) / two;
    if (n == 1) return f1;
    fint1 = gammaf(alpha + one) * gammaf(beta + two) / gammaf(apb + three);
    fint1 = fint1 * pow(two, apb + two);
    fint2 = gammaf(alpha + two) * gammaf(beta + two) / gammaf(apb + four);
    fint2 = fint2 * pow(two, apb + three);
    f2 = (two * (alpha + two) * fint1 - (apb + four) * fint2) * (apb + three) / four;
    if (n == 2) return f2;
    for (int i = 3; i <= n; i++)
    {
        double di = (double)(i - 1);
        double abn = alpha + beta + di;
        double abnn = abn + di;
        double a1 = -(two * (di + alpha) * (di + beta)) / (abn * abnn * (abnn + one));
        double a2 = (two * (alpha - beta)) / (abnn * (abnn + two));
        double a3 = (two * (abn + one)) / ((abnn + two) * (abnn + one));
        f3 = -(a2 * f2 + a1 * f1) / a3;
        f1 = f2;
        f2 = f3;
    }
    return f3;
}

/* ZWGLJD: Gauss-Lobatto-Jacobi points and weights. */
static void zwgljd(double *z, double *w, int np, double alpha, double beta)
{
    int n = np - 1;
    int nm1 = n - 1;
    double one = 1., two = 2.;
    double p, pd, pm1, pdm1, pm2, pdm2;
    if (nm1 > 0)
    {
        double alpg = alpha + one;
        double betg = beta + one;
        zwgjd(z + 1, w + 1, nm1, alpg, betg);
    }
    z[0] = -one;
    z[np - 1] = one;
    for (int i = 1; i < np - 1; i++) w[i] = w[i] / (one - z[i] * z[i]);
    jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, n, alpha, beta, z[0]);
    w[0] = endw1(n, alpha, beta) / (two * pd);
    jacobf(&p, &pd, &pm1, &pdm1, &pm2, &pdm2, n, alpha, beta, z[np - 1]);
    w[np - 1] = endw2(n, alpha, beta) / (two * pd);
}

/* PNLEG: Legendre polynomial of degree n at *z.  Writes *z = 0 when |*z| < 1e-25. */
static double pnleg(double *z, int n)
{
    if (fabs(*z) < 1.0e-25) *z = 0.0;
    double p1 = 1.;
    if (n == 0) return p1;
    double p2 = *z;
    double p3 = p2;
    for (int k = 1; k <= n - 1; k++)
    {
        double fk = (double)k;
        p3 = ((2. * fk + 1.) * (*z) * p2 - fk * p1) / (fk + 1.);
        p1 = p2;
        p2 = p3;
    }
    return p3;
}

/* PNDLEG: derivative of the Legendre polynomial of degree n at z. */
static double pndleg(double z, int n)
{
    double p1 = 1., p2 = z, p1d = 0., p2d = 1., p3d = 1., p3;
    for (int k = 1; k <= n - 1; k++)
    {
        double fk = (double)k;
        p3 = ((2. * fk + 1.) * z * p2 - fk * p1) / (fk + 1.);
        p3d = ((2. * fk + 1.) * p2 + (2. * fk + 1.) * z * p2d - fk * p1d) / (fk + 1.);
        p1 = p2;
        p2 = p3;
        p1d = p2d;
        p2d = p3d;
    }
    if (n == 0) return 0.;
    return p3d;
}

/* ---- the three entry points the reference declares in special_functions.hpp:10-12 ---- */

void oracle_zwgll(double *z, double *w, const int *np)
{
    zwgljd(z, w, *np, 0., 0.);
}

/*
 * DGLL(D, DT, Z, NZ, lzd): Fortran column-major.  D(I,J) lives at d[(I-1)+(J-1)*lzd],
 * DT(J,I) = D(I,J).  The reference calls dgll_(Dt_gll, D_gll, ...) and then reads D_gll
 * row-major, i.e. D_gll[i*n+j] = D(i+1,j+1) = dl_j/dxi(xi_i)  (domain.tpp:312-314).
 */
void oracle_dgll(double *d, double *dt, double *z, const int *nz_, const int *lzd_)
{
    int nz = *nz_, lzd = *lzd_;
    int n = nz - 1;
    if (nz == 1)
    {
        d[0] = 0.;
        return;
    }
    double fn = (double)n;
    double d0 = fn * (fn + 1.) / 4.;
    for (int i = 1; i <= nz; i++)
    {
        for (int j = 1; j <= nz; j++)
        {
            double v = 0.;
            if (i != j) v = pnleg(&z[i - 1], n) / (pnleg(&z[j - 1], n) * (z[i - 1] - z[j - 1]));
            if ((i == j) && (i == 1)) v = -d0;
            if ((i == j) && (i == nz)) v = d0;
            d[(i - 1) + (j - 1) * lzd] = v;
            dt[(j - 1) + (i - 1) * lzd] = v;
        }
    }
}

/* HGLL(I, Z, ZGLL, NZ): value at Z of the I-th (1-based) Lagrange interpolant through ZGLL. */
double oracle_hgll(const int *ii, double *z, double *zgll, const int *nz_)
{
    int i = *ii, nz = *nz_;
    double eps = 1.e-5;
    double dz = *z - zgll[i - 1];
    if (fabs(dz) < eps) return 1.;
    int n = nz - 1;
    double alfan = ((double)n) * (((double)n) + 1.);
    return -(1. - (*z) * (*z)) * pndleg(*z, n) / (alfan * pnleg(&zgll[i - 1], n) * (*z - zgll[i - 1]));
}
