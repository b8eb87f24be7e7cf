// mesh.cpp -- synthetic box-mesh writer in the reference's on-disk format (the Nek5000 dumps read at
// /root/reference/domain.tpp:43-224): per ladder degree N a directory <dir>/lx1_<N+1>/ with, per rank p,
// size_p.N.dat (text: dim n_x n_y n_z num_local_elements), x_/y_/z_ (f64), glo_num_ (i64),
// node_degree_ (i32), p_mask_ (f64), g_1..g_6_ (f64; order 11,22,33,12,13,23; 2D: 11,22,12,0,0,0).
//
// Conventions the reference relies on: vertex ids identical at every degree, so
// vertices are numbered first (subdomain.tpp:930-966); contiguous block partition, local elements in
// lexicographic order, global element id = rank-major (subdomain.tpp:219-280).  Geometric factors
// follow Nek5000: G_ab = w_i w_j w_k J (grad r_a . grad r_b), Jacobian by spectral differentiation.
#include "special_functions.hpp"
#include "../../../include/prfdd_b200.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace
{
struct Layout { int P[3]; };

Layout rank_layout(int nranks, int dimn)
{
    Layout L{{1, 1, 1}};
    int d = 0, r = nranks;
    while (r > 1)
    {
        if (r % 2) throw std::runtime_error("mesh: rank count must be a power of two");
        L.P[d % dimn] *= 2;
        r /= 2;
        d++;
    }
    return L;
}

template <typename T>
void dump(const std::string &dir, const char *name, int p, int N, const std::vector<T> &a)
{
    char fn[4096];
    snprintf(fn, sizeof(fn), "%s/%s_%d.%d.dat", dir.c_str(), name, p, N);
    FILE *f = fopen(fn, "wb");
    if (!f) throw std::runtime_error(std::string("mesh: cannot write ") + fn);
    fwrite(a.data(), sizeof(T), a.size(), f);
    fclose(f);
}
} // namespace

extern "C" int prfdd_mesh_generate_box_rank(const char *directory, int dimn, const int nel_in[3], int N, int nranks, double eps, int only_rank);

extern "C" int prfdd_mesh_generate_box(const char *directory, int dimn, const int nel_in[3], int N, int nranks, double eps)
{
    return prfdd_mesh_generate_box_rank(directory, dimn, nel_in, N, nranks, eps, -1);
}

extern "C" int prfdd_mesh_generate_box_rank(const char *directory, int dimn, const int nel_in[3], int N, int nranks, double eps, int only_rank)
{
    try
    {
        const int n = N + 1;
        int nel[3] = {nel_in[0], dimn >= 2 ? nel_in[1] : 1, dimn >= 3 ? nel_in[2] : 1};
        std::vector<double> z(n), w(n), D(n * n), Dt(n * n);
        zwgll_(z.data(), w.data(), &n);
        {
            std::vector<double> zz(z);
            dgll_(Dt.data(), D.data(), zz.data(), &n, &n); // D[i*n+m] = dl_m/dxi(xi_i)
        }
        Layout L = rank_layout(nranks, dimn);
        for (int d = 0; d < 3; d++)
            if (nel[d] % L.P[d]) throw std::runtime_error("mesh: elements per side not divisible by the rank layout");

        const long long gx = (long long)nel[0] * N + 1, gy = dimn >= 2 ? (long long)nel[1] * N + 1 : 1, gz = dimn >= 3 ? (long long)nel[2] * N + 1 : 1;
        const long long nvx = nel[0] + 1, nvy = nel[1] + 1;
        // id of a global grid node: vertices first, then every other node in lexicographic order
        const long long nv = (long long)(nel[0] + 1) * (dimn >= 2 ? nel[1] + 1 : 1) * (dimn >= 3 ? nel[2] + 1 : 1);
        auto is_v = [&](long long I, long long J, long long K) { return (I % N == 0) && (J % N == 0) && (K % N == 0); };
        // number of vertices with lexicographic index < (I,J,K) (K slowest): closed form
        auto verts_before = [&](long long I, long long J, long long K) {
            long long full_planes = (K + N - 1) / N;               // vertex planes strictly below K
            long long c = full_planes * nvx * (dimn >= 2 ? nvy : 1);
            if (K % N == 0)
            {
                long long full_rows = (J + N - 1) / N;             // vertex rows strictly below J in this plane
                c += full_rows * nvx;
                if (J % N == 0) c += (I + N - 1) / N;              // vertices strictly left of I in this row
            }
            return c;
        };
        auto node_id = [&](long long I, long long J, long long K) -> long long {
            if (is_v(I, J, K)) return 1 + I / N + (J / N) * nvx + (K / N) * nvx * nvy;
            long long lex = I + J * gx + K * gx * gy;              // nodes strictly before: lex
            long long nonv_before = lex - verts_before(I, J, K);
            return nv + nonv_before + 1;
        };
        auto mult1 = [&](long long ix, int ne) { return ((ix % N == 0) && ix > 0 && ix < (long long)ne * N) ? 2 : 1; };
        auto coord1 = [&](long long ix, int ne) {
            long long e = ix / N;
            if (e > ne - 1) e = ne - 1;
            long long l = ix - e * N;
            return ((double)e + 0.5 * (z[l] + 1.0)) / (double)ne;
        };
        const double pi = M_PI;
        auto deform = [&](double c[3]) {
            if (eps == 0.0) return;
            double s = 1.0;
            for (int d = 0; d < dimn; d++) s = s * std::sin(pi * c[d]);
            for (int d = 0; d < dimn; d++) c[d] = c[d] + eps * (0.5 + 0.25 * d) * s;
        };

        char sub[4096];
        snprintf(sub, sizeof(sub), "%s", directory);
        mkdir(sub, 0777);
        snprintf(sub, sizeof(sub), "%s/lx1_%d", directory, n);
        mkdir(sub, 0777);
        const std::string dir(sub);

        int npts = 1;
        for (int d = 0; d < dimn; d++) npts *= n;
        const int bl[3] = {nel[0] / L.P[0], nel[1] / L.P[1], nel[2] / L.P[2]};
        const int nk = dimn >= 3 ? n : 1, nj = dimn >= 2 ? n : 1;

        for (int p = 0; p < nranks; p++)
        {
            if (only_rank >= 0 && p != only_rank) continue;
            const int pc[3] = {p % L.P[0], (p / L.P[0]) % L.P[1], p / (L.P[0] * L.P[1])};
            const long long E = (long long)bl[0] * bl[1] * bl[2];
            std::vector<double> x(E * npts), y(E * npts), zc(E * npts), mask(E * npts);
            std::vector<long long> glo(E * npts);
            std::vector<int> deg(E * npts);
            std::vector<double> G[6];
            for (auto &g : G) g.assign(E * npts, 0.0);
            std::vector<double> X[3];
            for (auto &a : X) a.resize(npts);
            long long e = 0;
            for (int ez = 0; ez < bl[2]; ez++)
                for (int ey = 0; ey < bl[1]; ey++)
                    for (int ex = 0; ex < bl[0]; ex++, e++)
                    {
                        const long long g0[3] = {(long long)(pc[0] * bl[0] + ex) * N, (long long)(pc[1] * bl[1] + ey) * N, (long long)(pc[2] * bl[2] + ez) * N};
                        for (int k = 0; k < nk; k++)
                            for (int j = 0; j < nj; j++)
                                for (int i = 0; i < n; i++)
                                {
                                    const long long I = g0[0] + i, J = dimn >= 2 ? g0[1] + j : 0, K = dimn >= 3 ? g0[2] + k : 0;
                                    const int v = i + j * n + k * n * n;
                                    const long long o = e * npts + v;
                                    glo[o] = node_id(I, J, K);
                                    deg[o] = mult1(I, nel[0]) * (dimn >= 2 ? mult1(J, nel[1]) : 1) * (dimn >= 3 ? mult1(K, nel[2]) : 1);
                                    bool bd = (I == 0) || (I == gx - 1);
                                    if (dimn >= 2) bd = bd || (J == 0) || (J == gy - 1);
                                    if (dimn >= 3) bd = bd || (K == 0) || (K == gz - 1);
                                    mask[o] = bd ? 0.0 : 1.0;
                                    double c[3] = {coord1(I, nel[0]), dimn >= 2 ? coord1(J, nel[1]) : 0.0, dimn >= 3 ? coord1(K, nel[2]) : 0.0};
                                    deform(c);
                                    X[0][v] = c[0]; X[1][v] = c[1]; X[2][v] = c[2];
                                    x[o] = c[0]; y[o] = c[1]; zc[o] = c[2];
                                }
                        // geometric factors
                        for (int k = 0; k < nk; k++)
                            for (int j = 0; j < nj; j++)
                                for (int i = 0; i < n; i++)
                                {
                                    const int v = i + j * n + k * n * n;
                                    double d[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; // d[c][a] = d x_c / d r_a
                                    for (int m = 0; m < n; m++)
                                        for (int c = 0; c < dimn; c++)
                                        {
                                            d[c][0] += D[i * n + m] * X[c][m + j * n + k * n * n];
                                            if (dimn >= 2) d[c][1] += D[j * n + m] * X[c][i + m * n + k * n * n];
                                            if (dimn >= 3) d[c][2] += D[k * n + m] * X[c][i + j * n + m * n * n];
                                        }
                                    const long long o = e * npts + v;
                                    if (dimn == 2)
                                    {
                                        const double xr = d[0][0], xs = d[0][1], yr = d[1][0], ys = d[1][1];
                                        const double jac = xr * ys - xs * yr;
                                        const double rx = ys, ry = -xs, sx = -yr, sy = xr;
                                        const double W = w[j] * w[i];
                                        G[0][o] = W * (rx * rx + ry * ry) / jac;
                                        G[1][o] = W * (sx * sx + sy * sy) / jac;
                                        G[2][o] = W * (rx * sx + ry * sy) / jac;
                                    }
                                    else
                                    {
                                        const double xr = d[0][0], xs = d[0][1], xt = d[0][2];
                                        const double yr = d[1][0], ys = d[1][1], yt = d[1][2];
                                        const double zr = d[2][0], zs = d[2][1], zt = d[2][2];
                                        const double jac = xr * (ys * zt - yt * zs) - xs * (yr * zt - yt * zr) + xt * (yr * zs - ys * zr);
                                        const double rx = ys * zt - yt * zs, ry = xt * zs - xs * zt, rz = xs * yt - xt * ys;
                                        const double sx = yt * zr - yr * zt, sy = xr * zt - xt * zr, sz = xt * yr - xr * yt;
                                        const double tx = yr * zs - ys * zr, ty = xs * zr - xr * zs, tz = xr * ys - xs * yr;
                                        const double W = w[k] * w[j] * w[i];
                                        G[0][o] = W * (rx * rx + ry * ry + rz * rz) / jac;
                                        G[1][o] = W * (sx * sx + sy * sy + sz * sz) / jac;
                                        G[2][o] = W * (tx * tx + ty * ty + tz * tz) / jac;
                                        G[3][o] = W * (rx * sx + ry * sy + rz * sz) / jac;
                                        G[4][o] = W * (rx * tx + ry * ty + rz * tz) / jac;
                                        G[5][o] = W * (sx * tx + sy * ty + sz * tz) / jac;
                                    }
                                }
                    }
            char fn[4096];
            snprintf(fn, sizeof(fn), "%s/size_%d.%d.dat", dir.c_str(), p, N);
            FILE *f = fopen(fn, "w");
            if (!f) throw std::runtime_error(std::string("mesh: cannot write ") + fn);
            fprintf(f, "%d %d %d %d %lld\n", dimn, n, n, dimn == 3 ? n : 1, E);
            fclose(f);
            dump(dir, "x", p, N, x); dump(dir, "y", p, N, y); dump(dir, "z", p, N, zc);
            dump(dir, "glo_num", p, N, glo); dump(dir, "node_degree", p, N, deg); dump(dir, "p_mask", p, N, mask);
            for (int g = 0; g < 6; g++)
            {
                char nm[16];
                snprintf(nm, sizeof(nm), "g_%d", g + 1);
                dump(dir, nm, p, N, G[g]);
            }
        }
        return 0;
    }
    catch (const std::exception &ex)
    {
        fprintf(stderr, "prfdd_mesh_generate_box: %s\n", ex.what());
        return -1;
    }
}
