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
#include <stdexcept>
#include <tuple>
#include <typeinfo>
#include <vector>
#include "config.hpp"
#include "../../../include/prfdd_b200.h"

// fills the launch plan of a descriptor from the host copy of the row pointers (prfdd_csr_plan) and uploads the list of
// long rows, which must stay alive with the matrix
inline dev::memory plan_csr(prfdd_csr_matrix &desc, const int *ptr_hst)
{
    std::vector<int> rows((size_t)std::max(desc.num_rows / 20 + 1, 1));
    int count = prfdd_csr_plan(&desc, ptr_hst, rows.data(), (int)rows.size());
    if (count < 0) throw std::runtime_error("plan_csr: prfdd_csr_plan failed");
    dev::memory list;
    if (count > 0)
    {
        list = prfdd_host::device.malloc<int>(count);
        list.copyFrom(rows.data(), count * sizeof(int));
        desc.long_rows = list.as<int>();
        desc.num_long_rows = count;
    }
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
    dev::memory long_rows;         // rows given a whole warp
    prfdd_csr_matrix desc = {};    // device arrays + launch plan, as the C ABI takes them
    bool unit_values = false;      // every stored value is 1.0: the value stream is never read (Q, Q^T of a conforming region)
    bool one_entry_per_row = false; // row i holds exactly entry i: applied as an index map

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
        desc = prfdd_csr_matrix();
        desc.ptr = ptr.as<int>(); desc.col = col.as<int>(); desc.val = (const double *)val.ptr();
        desc.num_rows = num_rows;
        desc.num_cols = num_cols;
        long_rows = plan_csr(desc, ptr_hst.data());
        threads_per_row = desc.threads_per_row;
        static const bool no_unit = getenv("PRFDD_CSR_NO_UNIT") != nullptr;
        unit_values = !no_unit && typeid(DType) == typeid(double);
        for (int j = 0; j < num_nnz && unit_values; j++) unit_values = (val_hst[j] == (DType)1);
        one_entry_per_row = unit_values && num_nnz == num_rows;
        for (int i = 0; i <= num_rows && one_entry_per_row; i++) one_entry_per_row = (ptr_hst[i] == i);
        if (unit_values) desc.val = nullptr;
        if (one_entry_per_row) desc.ptr = nullptr;
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
        dev::check_rc(prfdd_csrm_multiply(Au.as<double>(), &desc, u.as<double>(), prfdd_host::device.stream), "CSR_Matrix::multiply");
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
        dev::check_rc(prfdd_csrm_multiply_weight(Au.as<double>(), &desc, u.as<double>(), weight.as<double>(), prfdd_host::device.stream), "CSR_Matrix::multiply_weight");
    }
};
