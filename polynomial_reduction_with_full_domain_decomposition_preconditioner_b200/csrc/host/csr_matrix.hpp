// csr_matrix.hpp -- CSR_Matrix<DType>: the reference's sparse type with the same surface
// (/root/reference/csr_matrix.hpp:15-56, csr_matrix.tpp): initialize, add_entry, assemble, print,
// multiply, multiply_range, multiply_weight, transpose, diagonal; public num_rows/num_cols/num_nnz
// and device arrays ptr/col/val.
//
// Semantics kept: entries with |v| <= sparse_tolerance (1e-12 double, 1e-6 float) are dropped AT
// INSERTION, before duplicates are summed (tpp:61-64, 79-80); COO is sorted by (row, col) and
// duplicates are summed (tpp:102-165) -- with a stable sort, so the sum order is insertion order;
// assemble() on an empty entry list returns without allocating (tpp:96).
// Differences: a host mirror of ptr/col/val is kept (the reference re-downloads the matrix for every
// transpose/diagonal/print), and the SpMV launch shape (lanes per row) is chosen from nnz/rows.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <tuple>
#include <typeinfo>
#include <vector>
#include "config.hpp"
#include "../../../include/prfdd_b200.h"

// rows much longer than the average (and than the lanes that walk them can absorb) go to the warp-per-row launch
// (prfdd_csr_set_long_rows); returns the device list that must stay alive with the matrix
inline dev::memory register_long_rows(const int *ptr_dev, const int *ptr_hst, int num_rows, double avg, int tpr)
{
    static const bool off = getenv("PRFDD_SPMV_NO_LONG_ROWS") != nullptr;
    const int threshold = std::max(std::max(16, 4 * tpr), (int)std::ceil(3.0 * avg));
    std::vector<int> rows;
    int longest = 0;
    for (int r = 0; r < num_rows && !off; r++)
    {
        const int len = ptr_hst[r + 1] - ptr_hst[r];
        longest = std::max(longest, len);
        if (len > threshold) rows.push_back(r);
    }
    dev::memory list;
    // worth a second launch only when some row is several times the threshold (the tail it removes is then longer than the launch)
    if (rows.empty() || tpr >= 32 || longest < 4 * threshold || (double)rows.size() > 0.05 * num_rows)
    {
        prfdd_csr_set_long_rows(ptr_dev, nullptr, 0, 0);
        return list;
    }
    list = prfdd_host::device.malloc<int>(rows.size());
    list.copyFrom(rows.data(), rows.size() * sizeof(int));
    prfdd_csr_set_long_rows(ptr_dev, list.as<int>(), (int)rows.size(), threshold);
    return list;
}

template <typename DType>
class CSR_Matrix
{
  private:
    int is_initialized = false;
    DType sparse_tolerance;
    std::vector<std::tuple<int, int, DType>> entries;

    void initialization_check()
    {
        if (not is_initialized)
        {
            printf("ERROR: CSR matrix has not been initialize\n");
            exit(EXIT_FAILURE);
        }
    }

  public:
    int num_rows = 0;
    int num_cols = 0;
    int num_nnz = 0;
    dev::memory ptr;
    dev::memory col;
    dev::memory val;

    // host mirror and launch shape
    std::vector<int> ptr_hst;
    std::vector<int> col_hst;
    std::vector<DType> val_hst;
    int threads_per_row = 1;
    dev::memory long_rows; // rows given a whole warp (prfdd_csr_set_long_rows)

    CSR_Matrix() {}
    CSR_Matrix(int num_rows_, int num_cols_) { initialize(num_rows_, num_cols_); }
    ~CSR_Matrix() {}

    void initialize(int num_rows_, int num_cols_)
    {
        num_rows = num_rows_;
        num_cols = num_cols_;
        num_nnz = 0;
        entries.clear();
        ptr_hst.clear(); col_hst.clear(); val_hst.clear();
        sparse_tolerance = (typeid(DType) == typeid(double)) ? 1.0e-12 : 1.0e-6;
        is_initialized = true;
    }

    void reserve(size_t n) { entries.reserve(n); }

    void add_entry(int row, int col_, DType val_)
    {
        if ((row < 0) or (row >= num_rows) or (col_ < 0) or (col_ >= num_cols))
        {
            printf("ERROR: Entry at (%d, %d) is outside the matrix of size (%d, %d)\n", row, col_, num_rows, num_cols);
            exit(EXIT_FAILURE);
        }
        if (std::abs(val_) > sparse_tolerance) entries.push_back(std::tuple<int, int, DType>(row, col_, val_));
    }

    void assemble()
    {
        if ((num_rows == 0) or (num_cols == 0) or (entries.size() == 0)) return;
        initialization_check();

        std::stable_sort(entries.begin(), entries.end(), [](const std::tuple<int, int, DType> &a, const std::tuple<int, int, DType> &b) {
            if (std::get<0>(a) != std::get<0>(b)) return std::get<0>(a) < std::get<0>(b);
            return std::get<1>(a) < std::get<1>(b);
        });

        ptr_hst.assign(num_rows + 1, 0);
        col_hst.clear();
        val_hst.clear();
        int cur_r = -1, cur_c = -1;
        for (auto &en : entries)
        {
            int r = std::get<0>(en), c = std::get<1>(en);
            if (r != cur_r or c != cur_c)
            {
                cur_r = r; cur_c = c;
                ptr_hst[r + 1]++;
                col_hst.push_back(c);
                val_hst.push_back(std::get<2>(en));
            }
            else
            {
                val_hst.back() += std::get<2>(en);
            }
        }
        for (int i = 1; i <= num_rows; i++) ptr_hst[i] += ptr_hst[i - 1];
        num_nnz = (int)col_hst.size();
        entries.clear();
        entries.shrink_to_fit();
        upload();
    }

    // adopt an already assembled host CSR (sorted columns)
    void set_csr(int num_rows_, int num_cols_, std::vector<int> &&p, std::vector<int> &&c, std::vector<DType> &&v)
    {
        initialize(num_rows_, num_cols_);
        ptr_hst = std::move(p); col_hst = std::move(c); val_hst = std::move(v);
        num_nnz = (int)col_hst.size();
        if (num_rows > 0 && num_cols > 0 && num_nnz > 0) upload();
    }

    void upload()
    {
        ptr = prfdd_host::device.malloc<int>(num_rows + 1);
        col = prfdd_host::device.malloc<int>(num_nnz);
        val = prfdd_host::device.malloc<DType>(num_nnz);
        ptr.copyFrom(ptr_hst.data(), (num_rows + 1) * sizeof(int));
        col.copyFrom(col_hst.data(), num_nnz * sizeof(int));
        val.copyFrom(val_hst.data(), num_nnz * sizeof(DType));
        double avg = (double)num_nnz / (double)std::max(num_rows, 1);
        threads_per_row = avg <= 10 ? 1 : avg <= 18 ? 2 : (avg > 60 && num_rows < 50000) ? 16 : 8; // measured on B200, profiles/r1_spmv_tpr.txt and r1_notes.txt
        long_rows = register_long_rows(ptr.as<int>(), ptr_hst.data(), num_rows, avg, threads_per_row);
    }

    void print(FILE *file_ptr = NULL, int offset = 0)
    {
        FILE *f = file_ptr ? file_ptr : prfdd_host::pstdout_file;
        if (!f) return;
        fprintf(f, "num_rows = %d, num_cols = %d, num_nnz = %d\n", num_rows, num_cols, num_nnz);
        if ((num_rows == 0) or (num_cols == 0) or (num_nnz == 0)) return;
        for (int i = 0; i < num_rows; i++)
            for (int j = ptr_hst[i]; j < ptr_hst[i + 1]; j++) fprintf(f, "(%d, %d): %.16g\n", i + offset, col_hst[j] + offset, val_hst[j]);
    }

    void transpose(CSR_Matrix &At)
    {
        At.initialize(num_cols, num_rows);
        if ((num_rows == 0) or (num_cols == 0)) return;
        At.reserve(num_nnz);
        for (int i = 0; i < num_rows; i++)
            for (int j = ptr_hst.empty() ? 0 : ptr_hst[i]; j < (ptr_hst.empty() ? 0 : ptr_hst[i + 1]); j++) At.add_entry(col_hst[j], i, val_hst[j]);
        At.assemble();
    }

    void diagonal(dev::memory D)
    {
        if ((num_rows == 0) or (num_cols == 0)) return;
        initialization_check();
        std::vector<DType> work(num_rows, 0);
        for (int i = 0; i < num_rows; i++)
            for (int j = ptr_hst[i]; j < ptr_hst[i + 1]; j++)
                if (i == col_hst[j]) { work[i] = val_hst[j]; break; }
        D.copyFrom(work.data(), num_rows * sizeof(DType));
    }

    void multiply(dev::memory &Au, dev::memory &u) { multiply((const dev::memory &)Au, (const dev::memory &)u); }
    void multiply(const dev::memory &Au, const dev::memory &u)
    {
        if ((num_rows == 0) or (num_cols == 0)) return;
        initialization_check();
        if (num_nnz == 0) { dev::check_rc(prfdd_set_to_value(Au.as<double>(), 0.0, num_rows, 0, prfdd_host::device.stream), "CSR_Matrix::multiply"); return; }
        dev::check_rc(prfdd_csr_multiply(Au.as<double>(), ptr.as<int>(), col.as<int>(), val.as<double>(), u.as<double>(), num_rows, threads_per_row, prfdd_host::device.stream), "CSR_Matrix::multiply");
    }

    void multiply_range(const dev::memory &Au, const dev::memory &u, int row_start, int row_end)
    {
        if ((num_rows == 0) or (num_cols == 0)) return;
        initialization_check();
        if (row_end < row_start)
        {
            printf("Row end (i_e = %d) has to be greater or equal to row start (i_s = %d)\n", row_end, row_start);
            exit(EXIT_FAILURE);
        }
        dev::check_rc(prfdd_csr_multiply_range(Au.as<double>(), ptr.as<int>(), col.as<int>(), val.as<double>(), u.as<double>(), row_start, row_end, threads_per_row, prfdd_host::device.stream), "CSR_Matrix::multiply_range");
    }

    void multiply_weight(const dev::memory &Au, const dev::memory &u, const dev::memory &weight)
    {
        if ((num_rows == 0) or (num_cols == 0)) return;
        initialization_check();
        if (num_nnz == 0) { dev::check_rc(prfdd_set_to_value(Au.as<double>(), 0.0, num_rows, 0, prfdd_host::device.stream), "CSR_Matrix::multiply_weight"); return; }
        dev::check_rc(prfdd_csr_multiply_weight(Au.as<double>(), ptr.as<int>(), col.as<int>(), val.as<double>(), u.as<double>(), weight.as<double>(), num_rows, threads_per_row, prfdd_host::device.stream), "CSR_Matrix::multiply_weight");
    }
};
