// region_helpers.hpp -- small helpers shared by the PR-FDD constructor: element corner / edge / face point lists in
// the reference's parameterisation (subdomain.tpp:1179-1494), a cache of the other ranks' mesh files, COO -> CSR.
#pragma once
#include <array>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include "amg.hpp"

namespace prfdd_multi
{
static const int EDGE_PAIRS_2D[4][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}};
static const int EDGE_PAIRS_3D[12][2] = {{0, 1}, {2, 3}, {0, 2}, {1, 3}, {4, 5}, {6, 7}, {4, 6}, {5, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
static const int FACE_QUADS[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 1, 4, 5}, {2, 3, 6, 7}, {0, 2, 4, 6}, {1, 3, 5, 7}};
// edges bounding a face in the order (bottom, top, left, right) of the face's own (a, b) parameterisation
static const int FACE_EDGES[6][4] = {{0, 1, 2, 3}, {4, 5, 6, 7}, {0, 4, 8, 9}, {1, 5, 10, 11}, {2, 6, 8, 10}, {3, 7, 9, 11}};

inline std::vector<int> corner_indices(int dimn, int n)
{
    std::vector<int> c = {0, n - 1, n * (n - 1), n * n - 1};
    if (dimn == 3)
        for (int q = 0; q < 4; q++) c.push_back(c[q] + n * n * (n - 1));
    return c;
}

// local point indices along edge eid (same parameterisation as subdomain.tpp:1197-1308)
inline std::vector<int> edge_points(int dimn, int n, int eid)
{
    std::vector<int> p(n);
    const int nn = n * n;
    for (int k = 0; k < n; k++)
    {
        if (dimn == 2)
        {
            const int v[4] = {k, k + (n - 1) * n, k * n, (n - 1) + k * n};
            p[k] = v[eid];
        }
        else
        {
            const int v[12] = {k, k + (n - 1) * n, k * n, (n - 1) + k * n,
                               k + (n - 1) * nn, k + (n - 1) * n + (n - 1) * nn, k * n + (n - 1) * nn, (n - 1) + k * n + (n - 1) * nn,
                               k * nn, (n - 1) + k * nn, (n - 1) * n + k * nn, (n - 1) + (n - 1) * n + k * nn};
            p[k] = v[eid];
        }
    }
    return p;
}

// local point indices of face fid as a flat n*n list, first index fastest (subdomain.tpp:1366-1431)
inline std::vector<int> face_points(int n, int fid)
{
    std::vector<int> p(n * n);
    const int nn = n * n;
    for (int b = 0; b < n; b++)
        for (int a = 0; a < n; a++)
        {
            int v = 0;
            switch (fid)
            {
            case 0: v = a + b * n; break;
            case 1: v = a + b * n + (n - 1) * nn; break;
            case 2: v = a + b * nn; break;
            case 3: v = a + (n - 1) * n + b * nn; break;
            case 4: v = a * n + b * nn; break;
            default: v = (n - 1) + a * n + b * nn; break;
            }
            p[a + b * n] = v;
        }
    return p;
}

// whole-file cache of the other ranks' mesh arrays
template <typename T>
class FileCache
{
    std::map<std::tuple<std::string, int, int>, std::vector<T>> cache;
    std::string dir;

  public:
    explicit FileCache(const std::string &d) : dir(d) {}
    const std::vector<T> &get(const char *name, int rank, int degree, size_t count)
    {
        auto key = std::make_tuple(std::string(name), rank, degree);
        auto it = cache.find(key);
        if (it != cache.end()) return it->second;
        char fn[4096];
        snprintf(fn, sizeof(fn), "%s/lx1_%d/%s_%d.%d.dat", dir.c_str(), degree + 1, name, rank, degree);
        std::vector<T> v(count);
        FILE *f = fopen(fn, "rb");
        if (!f || fread(v.data(), sizeof(T), count, f) != count) throw std::runtime_error(std::string("Subdomain: cannot read ") + fn);
        fclose(f);
        return cache.emplace(key, std::move(v)).first->second;
    }
    void clear() { cache.clear(); }
};

inline amg::HostCSR csr_from_coo(int nr, int nc, std::vector<std::tuple<int, int, double>> &e)
{
    std::stable_sort(e.begin(), e.end(), [](const std::tuple<int, int, double> &a, const std::tuple<int, int, double> &b) {
        if (std::get<0>(a) != std::get<0>(b)) return std::get<0>(a) < std::get<0>(b);
        return std::get<1>(a) < std::get<1>(b);
    });
    amg::HostCSR M;
    M.num_rows = nr;
    M.num_cols = nc;
    M.ptr.assign(nr + 1, 0);
    int cr = -1, cc = -1;
    for (auto &t : e)
    {
        if (std::get<0>(t) != cr || std::get<1>(t) != cc)
        {
            cr = std::get<0>(t); cc = std::get<1>(t);
            M.ptr[cr + 1]++;
            M.col.push_back(cc);
            M.val.push_back(std::get<2>(t));
        }
        else
            M.val.back() += std::get<2>(t);
    }
    for (int i = 0; i < nr; i++) M.ptr[i + 1] += M.ptr[i];
    return M;
}
} // namespace prfdd_multi
