// globals.cpp -- definitions of the process-wide globals of config.hpp and the small library-level
// C ABI (version, device selection, buffers).
#include <cstdio>
#include "config.hpp"
#include "../../../include/prfdd_b200.h"
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

namespace prfdd_host
{
int dim = 0;
int proc_id = 0;
int num_procs = 1;
int verbose = 0;
dev::device device;
Timer<double> timer;
Comm comm_world;
FILE *pstdout_file = nullptr;
} // namespace prfdd_host

extern "C" {

const char *prfdd_version(void) { return "prfdd_b200 0.1 (sm_100a)"; }

const char *prfdd_error_string(int code)
{
    switch (code)
    {
    case 0: return "success";
    case -1: return "exception in host code (see stderr)";
    case -2: return "missing reduction workspace";
    case -3: return "too many vectors in a fused operation (max 32)";
    case -4: return "unsupported number of GLL points (2..16)";
    case -5: return "could not read the derivative matrix from device memory";
    case -6: return "threads_per_row must be one of 1,2,4,8,16,32";
    case -7: return "row_end < row_start";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown prfdd error";
    }
}

int prfdd_halo_build_lists(int me, int nprocs, const long long *ids, const long long *offsets, int *peers, int *peer_count, int *peer_offset, int *idx,
                           long long idx_capacity, long long *total)
{
    std::unordered_map<long long, int> mine;
    for (long long i = offsets[me]; i < offsets[me + 1]; i++) mine[ids[i]] = (int)(i - offsets[me]);
    int np = 0;
    long long n = 0;
    for (int p = 0; p < nprocs; p++)
    {
        if (p == me) continue;
        std::vector<std::pair<long long, int>> shared;
        for (long long i = offsets[p]; i < offsets[p + 1]; i++)
        {
            auto it = mine.find(ids[i]);
            if (it != mine.end()) shared.push_back({ids[i], it->second});
        }
        if (shared.empty()) continue;
        std::sort(shared.begin(), shared.end());
        if (n + (long long)shared.size() > idx_capacity) return -8;
        peers[np] = p;
        peer_offset[np] = (int)n;
        peer_count[np] = (int)shared.size();
        for (auto &s : shared) idx[n++] = s.second;
        np++;
    }
    *total = n;
    return np;
}

int prfdd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int prfdd_set_device(int d) { return (int)cudaSetDevice(d); }
int prfdd_malloc(void **dptr, size_t bytes) { return (int)cudaMalloc(dptr, bytes); }
int prfdd_free(void *dptr) { return (int)cudaFree(dptr); }
int prfdd_memcpy_h2d(void *dst, const void *src, size_t bytes, prfdd_stream_t s) { return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)s); }
int prfdd_memcpy_d2h(void *dst, const void *src, size_t bytes, prfdd_stream_t s) { return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s); }
int prfdd_memcpy_d2d(void *dst, const void *src, size_t bytes, prfdd_stream_t s) { return (int)cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s); }
int prfdd_stream_synchronize(prfdd_stream_t s) { return (int)cudaStreamSynchronize((cudaStream_t)s); }

// Field output (Domain::output, domain.tpp:373-524; Subdomain::output, subdomain.tpp:4648-4791): the reference writes one Silo
// file (UCD mesh of the low-order cells, every GLL cell a quad / hexahedron, node-centred fields).  Silo is not available; the
// same mesh and fields go to a legacy-VTK unstructured grid (cell types 9 / 12 have the reference's vertex order), readable by
// VisIt and ParaView.  Host arrays, element-major points.
int prfdd_write_vtk(const char *path, int dim, int n, int num_elements, const double *x, const double *y, const double *z, int num_fields,
                    const char *const *field_names, const double *const *fields)
{
    if (n < 2 || num_elements < 0) return -8;
    std::vector<int> ns((size_t)std::max(num_elements, 1), n);
    return prfdd_write_vtk_mixed(path, dim, num_elements, ns.data(), x, y, z, num_fields, field_names, fields);
}

// the same for elements of different degrees (a subdomain region: own elements at N, rings at the ladder degrees, extended N = 1
// elements; Subdomain::output, subdomain.tpp:4648-4791): element e has n_of_element[e] points per side, points element-major
int prfdd_write_vtk_mixed(const char *path, int dim, int num_elements, const int *n_of_element, const double *x, const double *y, const double *z,
                          int num_fields, const char *const *field_names, const double *const *fields)
{
    if (!path || (dim != 2 && dim != 3) || num_elements < 0 || (num_elements > 0 && !n_of_element) || !x || !y || (dim == 3 && !z)) return -8;
    long long num_points = 0, num_cells = 0;
    for (int e = 0; e < num_elements; e++)
    {
        const long long n = n_of_element[e];
        if (n < 2) return -8;
        num_points += (dim == 2) ? n * n : n * n * n;
        num_cells += (dim == 2) ? (n - 1) * (n - 1) : (n - 1) * (n - 1) * (n - 1);
    }
    FILE *f = fopen(path, "w");
    if (!f) return -2;
    const int nv = (dim == 2) ? 4 : 8;
    fprintf(f, "# vtk DataFile Version 3.0\nField data\nASCII\nDATASET UNSTRUCTURED_GRID\nPOINTS %lld double\n", num_points);
    for (long long p = 0; p < num_points; p++) fprintf(f, "%.17g %.17g %.17g\n", x[p], y[p], dim == 3 ? z[p] : 0.0);
    fprintf(f, "CELLS %lld %lld\n", num_cells, num_cells * (nv + 1));
    long long element_offset = 0;
    for (int e = 0; e < num_elements; e++)
    {
        const int n = n_of_element[e];
        for (int sz = 0; sz < (dim == 3 ? n - 1 : 1); sz++)
            for (int sy = 0; sy < n - 1; sy++)
                for (int sx = 0; sx < n - 1; sx++)
                {
                    const long long b = element_offset + sx + (long long)sy * n + (long long)sz * n * n;
                    if (dim == 2) fprintf(f, "4 %lld %lld %lld %lld\n", b, b + 1, b + 1 + n, b + n);
                    else
                    {
                        const long long t = b + (long long)n * n;
                        fprintf(f, "8 %lld %lld %lld %lld %lld %lld %lld %lld\n", b, b + 1, b + 1 + n, b + n, t, t + 1, t + 1 + n, t + n);
                    }
                }
        element_offset += (dim == 2) ? (long long)n * n : (long long)n * n * n;
    }
    fprintf(f, "CELL_TYPES %lld\n", num_cells);
    for (long long c = 0; c < num_cells; c++) fprintf(f, dim == 2 ? "9\n" : "12\n");
    if (num_fields > 0) fprintf(f, "POINT_DATA %lld\n", num_points);
    for (int k = 0; k < num_fields; k++)
    {
        fprintf(f, "SCALARS %s double 1\nLOOKUP_TABLE default\n", field_names[k]);
        for (long long p = 0; p < num_points; p++) fprintf(f, "%.17g\n", fields[k][p]);
    }
    return fclose(f) == 0 ? 0 : -2;
}

} // extern "C"
