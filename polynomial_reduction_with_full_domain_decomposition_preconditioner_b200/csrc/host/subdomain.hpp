// subdomain.hpp -- Subdomain<DType>: the PR-FDD preconditioner (polynomial reduction + full domain
// decomposition), same class surface as the reference (/root/reference/subdomain.hpp:72-251):
// Subdomain(domains, N, reduction, sub_overlap, super_overlap); stiffness_matrix,
// direct_stiffness_summation, flexible_conjugate_gradient(u_l, f_l), generalized_minimum_residual(u_l, f_l);
// same solver defaults (subdomain.hpp:228-238).
//
// One preconditioner application is a rank-local Krylov solve on the composite problem
//   [ own elements at degree N | overlap rings at the ladder degrees | rest of the mesh at N=1, AMG-coarsened ]
// preconditioned by a Chebyshev-smoothed AMG V-cycle on a P1-simplex FEM discretisation.
//
// B200-first differences from the reference:
//   * the whole application -- tree operator, 4 Arnoldi steps, V-cycles, Gram-Schmidt, Givens, back
//     substitution -- runs on one stream with NO host synchronisation: all scalars live in a device-side
//     prfdd_krylov_state, so the application is captured once in a CUDA graph and replayed
//     (the reference synchronises j+3 times per Arnoldi step and runs the coarse AMG levels on the host);
//   * the variable-degree operator runs per run of equal-degree elements with the fused kernel instead of
//     per-point level/offset/vertex lookups (subdomain.okl:4-101);
//   * Qt*w-assembled copies of the Krylov basis are cached, so the j+1 Gram-Schmidt dots of a column are one
//     gather + one fused multi-dot instead of 2(j+1) SpMVs + (j+1) reductions (subdomain.tpp:4277-4307);
//   * the restrictions of the ladder are one fused kernel per step (subdomain.okl:284-366).
//
// Status: num_procs == 1 (empty superdomain) is complete; the multi-rank region / superdomain
// construction is in subdomain_multi.hpp.
#pragma once
#include <thread>
#include <atomic>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "config.hpp"
#include "domain.hpp"
#include "csr_matrix.hpp"
#include "math.hpp"
#include "special_functions.hpp"
#include "amg.hpp"
#include "region_helpers.hpp"
#include "../../../include/prfdd_b200.h"

template <typename DType>
struct Stiffness_Operator
{
    int num_dofs = 0;
    int num_points = 0;
    int num_extended_dofs = 0;

    CSR_Matrix<DType> Q;
    CSR_Matrix<DType> Qt;

    CSR_Matrix<DType> A;
    CSR_Matrix<DType> P;
    CSR_Matrix<DType> Pt;

    std::vector<dev::memory> D_hat;
    dev::memory geom_fact[NUM_GEOM_FACTS];

    // runs of equal-degree elements (replace the per-point element/vertex/level/offset arrays, tpp:1603-1630)
    std::vector<int> bucket_first_point, bucket_num_elements, bucket_n;
    std::vector<const double *> bucket_D, bucket_D_hst; // device / host copies of the run's derivative matrix
};

template <typename DType>
class Subdomain
{
    using memory = dev::memory;

  public:
    struct Level
    {
        int num_points = 0;
        int num_elements = 0;
        int poly_degree = 0;
        int offset = 0;
    };

  private:
    // Work arrays
    std::vector<std::vector<DType>> work_hst;
    std::vector<memory> work_dev;

    // Geometry
    int poly_reduction = 0;
    int subdomain_overlap = 1;
    int superdomain_overlap = 1;
    std::vector<int> poly_degree;
    int num_levels = 0;
    std::vector<Level> levels;

    // Coarse to fine interpolator
    std::map<std::pair<int, int>, std::pair<std::vector<DType>, memory>> J_cf;

    // Reference operator
    std::vector<std::pair<std::vector<DType>, memory>> D_hat;
    std::vector<std::vector<double>> r_gll;

    // Subdomain operator
    int num_subdomain_elems = 0;
    int num_subdomain_points = 0;
    int num_subdomain_extended_elems = 0;
    int num_subdomain_extended_points = 0;
    Stiffness_Operator<DType> subdomain_operator;
    std::vector<Element<DType>> subdomain_region;

    // Superdomain operator
    CSR_Matrix<DType> Qt_coarse;
    int num_superdomain_elems = 0;
    int num_superdomain_extended_elems = 0;
    Stiffness_Operator<DType> superdomain_operator;
    std::vector<Element<DType>> superdomain_region;

    // multi-rank bookkeeping (subdomain_multi.hpp)
    std::vector<int> proc_count, proc_offset;
    int num_total_elements = 0;
    int num_comp_levels = 1;
    std::vector<long long> dof_num_coarse; // global N = 1 dof of every (element, vertex)
    std::vector<int> dof_marker, dof_sup;
    std::map<std::pair<int, int>, std::vector<DType>> J_cf_fem; // hat-function interpolation for hanging nodes (tpp:2754-2783)
    struct TreeExchange
    {
        std::vector<int> peers, send_count, recv_count, send_offset, recv_offset;
        int send_total = 0, recv_total = 0, coarse_per_rank = 0;
        memory send_idx, recv_idx, send_buf, recv_buf, coarse_all, coarse_dofs, level_buf;
    } tree;

    // Interface assembly
    int num_interface_dofs = 0;
    CSR_Matrix<DType> Q_int;
    CSR_Matrix<DType> Qt_int;
    CSR_Matrix<DType> QQt_int;
    bool interface_is_identity = true;

    // Preconditioner
    amg::Hierarchy amg_fem;
    amg::HostCSR A_fem_hst;

    // Solver
    int num_dofs = 0;
    memory norm_weight;
    memory inner_weight;
    std::vector<DType> norm_weight_hst, inner_weight_hst;

    memory f, u_k, r_k, r_kp1, q_k, z_k, p_k;
    std::vector<memory> V, Z, aV;
    memory aq;
    memory kstate; // prfdd_krylov_state on the device
    prfdd_reduce_ws *ws = nullptr;

    Math<DType> math;
    prfdd_options opt;
    int own_points = 0; // levels[0].num_points

    // CUDA graph of one preconditioner application, keyed by (input, output) pointers
    struct GraphKey { const void *in; void *out; int type; bool operator<(const GraphKey &o) const { return std::tie(in, out, type) < std::tie(o.in, o.out, o.type); } };
    std::map<GraphKey, cudaGraphExec_t> graphs;
    long long launches_per_apply = 0;
    double bytes_per_apply = 0.0; // algorithmic bytes of one application (prfdd_algorithmic_bytes)

    cudaStream_t st() const { return prfdd_host::device.stream; }
    static double *dp(const memory &m) { return m.as<double>(); }
    prfdd_krylov_state *ks() const { return kstate.as<prfdd_krylov_state>(); }
    double *ks_ptr(size_t off) const { return reinterpret_cast<double *>(reinterpret_cast<char *>(kstate.ptr()) + off); }

    void build_single_rank(std::map<int, std::unique_ptr<Domain<DType>>> &domains);
    void build_multi_rank(std::map<int, std::unique_ptr<Domain<DType>>> &domains);
    void build_region_Q(std::vector<Element<DType>> &region, CSR_Matrix<DType> &Q);
    void build_superdomain(amg::Hierarchy &H, const std::vector<int> &sub_ids, const std::vector<int> &sup_ids, const std::unordered_set<long long> &interface_glo_num,
                           const std::vector<long long> &glo_num_coarse, int num_coarse_dofs);
    void build_interface(const std::vector<int> &sub_ids, const std::vector<int> &sup_ids, const std::vector<int> &subdomain_partition, int n_interface);
    void setup_tree_exchange();
    void tree_operator_multi(const memory &Tu, const memory &u);
    void q1_element_matrix(const DType *const g[NUM_GEOM_FACTS], std::vector<DType> &Ae);
    std::pair<std::vector<int>, std::vector<int>> matching(const Element<DType> &ei, const Element<DType> &ej, int kind, int idx);
    int min_degree_edge_neighbor(const std::vector<Element<DType>> &region, const Element<DType> &el, int eid);
    void assemble_low_order_fem();
    void allocate_solver();
    void ranking(std::vector<DType> &data, int size);

    void tree_operator(const memory &Tu, const memory &u);
    void low_order_preconditioner(const memory &z, const memory &r, const memory *r_assembled = nullptr);
    void assemble_weighted(const memory &dst, const memory &src);
    void assemble(const memory &dst, const memory &src);
    void residual_norm_dev(const memory &r, double *out);
    void gmres_body(const memory &u_l, const memory &f_l);
    void fcg_body(const memory &u_l, const memory &f_l);
    template <class Body>
    void run_captured(int type, const memory &u_l, const memory &f_l, Body body);

  public:
    const char *data_type = (typeid(DType) == typeid(double)) ? "double" : "float";

    Subdomain() {}
    Subdomain(std::map<int, std::unique_ptr<Domain<DType>>> &domains, int poly_degree_, int poly_reduction_, int subdomain_overlap_, int superdomain_overlap_, const prfdd_options &opt_);
    ~Subdomain()
    {
        for (auto &g : graphs) cudaGraphExecDestroy(g.second);
        if (ws) prfdd_reduce_ws_destroy(ws);
    }
    Subdomain(const Subdomain &) = delete;
    Subdomain &operator=(const Subdomain &) = delete;

    // Solver
    int num_iterations = 0;
    int num_vectors = 4;
    int max_iterations = 4;
    bool use_preconditioner = true;
    DType tolerance = (typeid(DType) == typeid(double)) ? 1.0e-12 : 1.0e-06;
    DType epsilon = (typeid(DType) == typeid(double)) ? 1.0e-12 : 1.0e-06;

    // Preconditioner
    int num_vcycles = 1;
    int cheby_order = 2;
    int level_cutoff = 5; // kept for the surface; every level runs on the GPU here

    // Elements
    int num_values = 0;
    std::vector<Element<DType>> elements;

    // Member functions
    void direct_stiffness_summation(const memory &QQtu, const memory &u);
    void stiffness_matrix(const memory &Au, const memory &u);
    void flexible_conjugate_gradient(memory &u_l, memory &f_l, bool print_history = true, bool use_relative = false);
    void generalized_minimum_residual(memory &u_l, memory &f_l, bool print_history = true, bool use_relative = false);

    // subdomain.tpp:4648-4791: region mesh (elements of the ladder degrees as low-order cells) + node fields, legacy VTK instead of Silo
    int output(const std::string &output_name, const std::vector<std::pair<std::string, const memory *>> &fields);
    int output_last_application(const std::string &output_name) { return output(output_name, {{"f", &f}, {"u", &u_k}}); } // composite right-hand side and solution of the last inner solve

    // C ABI support
    long long query(int what);
    long long get_array(int what, void *dst, long long cap);
    int apply(int what, const double *in_host, double *out_host);
    int time_spmv(int reps, double out[6]);
    int profile_vcycle(int reps, char *text, int cap);
};

// ---------------------------------------------------------------------------------------------
// construction
// ---------------------------------------------------------------------------------------------
template <typename DType>
Subdomain<DType>::Subdomain(std::map<int, std::unique_ptr<Domain<DType>>> &domains, int poly_degree_, int poly_reduction_, int subdomain_overlap_, int superdomain_overlap_, const prfdd_options &opt_)
{
    using namespace prfdd_host;
    opt = opt_;
    num_vectors = opt.inner_num_vectors;
    max_iterations = opt.inner_max_iterations;
    tolerance = opt.inner_tolerance;
    num_vcycles = opt.num_vcycles;
    cheby_order = opt.cheby_order;
    if (num_vectors > PRFDD_KRYLOV_MAXV) throw std::runtime_error("Subdomain: num_vectors > PRFDD_KRYLOV_MAXV");

    // Construct levels (subdomain.tpp:93-120)
    poly_reduction = poly_reduction_;
    subdomain_overlap = subdomain_overlap_;
    superdomain_overlap = superdomain_overlap_;
    poly_degree.push_back(poly_degree_);
    while (poly_degree.back() > 1)
    {
        int reduced = poly_degree.back() - poly_reduction;
        poly_degree.push_back(reduced >= 1 ? reduced : 1);
    }
    num_levels = (int)poly_degree.size();
    levels.resize(num_levels);
    for (int l = 0; l < num_levels; l++)
    {
        Domain<DType> &d = *domains.at(poly_degree[l]);
        levels[l].num_points = d.num_local_points;
        levels[l].num_elements = d.num_local_elements;
        levels[l].poly_degree = d.poly_degree;
        if (l > 0) levels[l].offset = levels[l - 1].offset + levels[l - 1].num_points;
    }
    own_points = levels[0].num_points;

    // Prolongation/restriction reference operators (tpp:129-164)
    r_gll.resize(num_levels);
    for (int l = 0; l < num_levels; l++)
    {
        int n_l = poly_degree[l] + 1;
        std::vector<double> w_gll(n_l);
        r_gll[l].resize(n_l);
        zwgll_(r_gll[l].data(), w_gll.data(), &n_l);
    }
    for (int l_f = 0; l_f < num_levels - 1; l_f++)
        for (int l_c = l_f + 1; l_c < num_levels; l_c++)
        {
            int n_f = poly_degree[l_f] + 1, n_c = poly_degree[l_c] + 1;
            std::pair<int, int> idx(poly_degree[l_c], poly_degree[l_f]);
            J_cf[idx].first.resize(n_c * n_f);
            for (int i = 0; i < n_f; i++)
                for (int j = 1; j <= n_c; j++) J_cf[idx].first[i * n_c + (j - 1)] = (DType)(hgll_(&j, &r_gll[l_f][i], r_gll[l_c].data(), &n_c));
            J_cf[idx].second = device.malloc<DType>(n_c * n_f);
            J_cf[idx].second.copyFrom(J_cf[idx].first.data(), n_c * n_f * sizeof(DType));
        }

    // Operator (tpp:166-196)
    D_hat.resize(num_levels);
    for (int l = 0; l < num_levels; l++)
    {
        int n_l = poly_degree[l] + 1;
        std::vector<double> D_gll(n_l * n_l), Dt_gll(n_l * n_l);
        dgll_(Dt_gll.data(), D_gll.data(), r_gll[l].data(), &n_l, &n_l);
        D_hat[l].first.assign(D_gll.begin(), D_gll.end());
        D_hat[l].second = device.malloc<DType>(n_l * n_l);
        D_hat[l].second.copyFrom(D_hat[l].first.data(), n_l * n_l * sizeof(DType));
        subdomain_operator.D_hat.push_back(D_hat[l].second);
        superdomain_operator.D_hat.push_back(D_hat[l].second);
    }

    // hat-function interpolation between the GLL grids of two levels, for hanging nodes of the low-order FEM
    // (tpp:2754-2783; r_gll as dgll_ left it)
    for (int l_f = 0; l_f < num_levels - 1; l_f++)
        for (int l_c = l_f + 1; l_c < num_levels; l_c++)
        {
            const int N_f = poly_degree[l_f], N_c = poly_degree[l_c], n_f = N_f + 1, n_c = N_c + 1;
            std::vector<DType> &J = J_cf_fem[std::pair<int, int>(N_c, N_f)];
            J.assign(n_c * n_f, 0.0);
            J[0] = 1.0;
            for (int i = 1; i < N_f; i++)
                for (int j = 0; j < N_c; j++)
                    if ((r_gll[l_c][j] <= r_gll[l_f][i]) and (r_gll[l_f][i] <= r_gll[l_c][j + 1]))
                    {
                        J[i * n_c + (j + 0)] = (r_gll[l_c][j + 1] - r_gll[l_f][i]) / (r_gll[l_c][j + 1] - r_gll[l_c][j]);
                        J[i * n_c + (j + 1)] = (r_gll[l_f][i + 0] - r_gll[l_c][j]) / (r_gll[l_c][j + 1] - r_gll[l_c][j]);
                    }
            J[(n_f - 1) * n_c + (n_c - 1)] = 1.0;
        }

    dev::check_rc(prfdd_reduce_ws_create(&ws), "prfdd_reduce_ws_create");

    if (num_procs == 1)
        build_single_rank(domains);
    else
        build_multi_rank(domains);

    allocate_solver();
}

// dense ranking of subdomain.tpp:881-918: equal values -> equal ranks, the value 0 -> rank 0
template <typename DType>
void Subdomain<DType>::ranking(std::vector<DType> &data, int size)
{
    if (size == 0) return;
    std::vector<std::pair<unsigned int, DType>> entries(size);
    for (int i = 0; i < size; i++) { entries[i].first = i; entries[i].second = data[i]; }
    std::sort(entries.begin(), entries.end(), [](const std::pair<unsigned int, DType> &a, const std::pair<unsigned int, DType> &b) { return a.second < b.second; });
    DType value = entries[0].second;
    DType rank = (value == 0.0) ? 0.0 : 1.0;
    entries[0].second = rank;
    for (int i = 1; i < size; i++)
    {
        auto &entry = entries[i];
        if (entry.second == value)
            entry.second = rank;
        else
        {
            rank += 1.0;
            value = entry.second;
            entry.second = rank;
        }
    }
    for (int i = 0; i < size; i++) data[entries[i].first] = entries[i].second;
}

template <typename DType>
void Subdomain<DType>::build_single_rank(std::map<int, std::unique_ptr<Domain<DType>>> &domains)
{
    using namespace prfdd_host;
    Domain<DType> &domain = *domains.at(poly_degree[0]);
    const int num_local_elements = domain.num_local_elements;

    // Construct computational regions (tpp:455-579): own elements at degree N; nothing else exists
    subdomain_region.reserve(num_local_elements);
    for (int e = 0; e < num_local_elements; e++)
    {
        subdomain_region.push_back(domain.elements[e]); // id = proc_offset (0) + e, all fields pulled from the owner (tpp:644-805)
        num_subdomain_elems++;
        num_subdomain_extended_elems++;
    }
    for (auto &elem : subdomain_region)
    {
        num_subdomain_points += elem.num_points;
        num_subdomain_extended_points += elem.num_points;
        for (int v = 0; v < elem.num_points; v++) { elem.loc_num[v] = elem.offset + v; elem.dof_num[v] = 0; }
    }
    for (int g = 0; g < NUM_GEOM_FACTS; g++) subdomain_operator.geom_fact[g] = domain.geom_fact[g]; // aliases the Domain's device arrays

    for (int e = 0; e < num_subdomain_elems; e++) elements.push_back(subdomain_region[e]);

    // Global numbering (tpp:920-1176): level-0 offset is 0, no non-conforming entities, no interface, no extended nodes
    {
        const int np = num_subdomain_extended_points;
        std::vector<DType> w(np);
        for (auto &elem : subdomain_region)
            for (int v = 0; v < elem.num_points; v++) w[elem.offset + v] = (DType)(elem.glo_num[v]);
        ranking(w, np);
        for (auto &elem : subdomain_region)
            for (int v = 0; v < elem.num_points; v++) elem.glo_num[v] = (long long)(w[elem.offset + v]);
        for (auto &elem : subdomain_region)
            for (int v = 0; v < elem.num_points; v++) w[elem.offset + v] = (DType)(elem.glo_num[v]) * elem.dirichlet_mask[v];
        ranking(w, np);
        for (auto &elem : subdomain_region)
            for (int v = 0; v < elem.num_points; v++) elem.dof_num[v] = (long long)(w[elem.offset + v]);
    }

    // Region operator setup (tpp:1496-1585): conforming region -> only the "vertices" loop contributes
    {
        auto &Q = subdomain_operator.Q;
        int num_points = subdomain_region.empty() ? 0 : subdomain_region.back().offset + subdomain_region.back().num_points;
        int ndofs = 0;
        for (auto &elem : subdomain_region)
            for (auto dof : elem.dof_num) ndofs = std::max(ndofs, (int)dof);
        Q.initialize(num_points, ndofs);
        Q.reserve(num_points);
        for (auto &elem_i : subdomain_region)
            for (int vid = 0; vid < elem_i.num_points; vid++)
                if (elem_i.dof_num[vid] > 0) Q.add_entry(elem_i.loc_num[vid], (int)elem_i.dof_num[vid] - 1, 1.0);
        Q.assemble();
        subdomain_operator.Q.transpose(subdomain_operator.Qt);
    }

    // Subdomain stiffness operator setup (tpp:1587-1630)
    subdomain_operator.num_dofs = 0;
    for (int e = 0; e < num_subdomain_elems; e++)
        subdomain_operator.num_dofs = std::max(subdomain_operator.num_dofs, (int)(*std::max_element(subdomain_region[e].dof_num.begin(), subdomain_region[e].dof_num.end())));
    subdomain_operator.num_points = subdomain_operator.Q.num_rows;
    subdomain_operator.num_extended_dofs = subdomain_operator.Q.num_cols;
    {
        // runs of equal degree, in region order
        std::unordered_map<int, int> level_degree;
        for (int l = 0; l < num_levels; l++) level_degree[poly_degree[l]] = l;
        size_t e = 0;
        while (e < subdomain_region.size())
        {
            size_t e2 = e;
            while (e2 < subdomain_region.size() && subdomain_region[e2].poly_degree == subdomain_region[e].poly_degree) e2++;
            subdomain_operator.bucket_first_point.push_back(subdomain_region[e].offset);
            subdomain_operator.bucket_num_elements.push_back((int)(e2 - e));
            subdomain_operator.bucket_n.push_back(subdomain_region[e].poly_degree + 1);
            subdomain_operator.bucket_D.push_back(dp(D_hat[level_degree[subdomain_region[e].poly_degree]].second));
            subdomain_operator.bucket_D_hst.push_back(D_hat[level_degree[subdomain_region[e].poly_degree]].first.data());
            e = e2;
        }
    }

    // Superdomain: empty.  The reference would hand HYPRE zero-sized matrices here (tpp:2426-2431).
    superdomain_operator.num_dofs = 0;
    superdomain_operator.num_extended_dofs = 0;
    superdomain_operator.num_points = 0;

    // Interface operator (tpp:2581-2729): identities
    num_interface_dofs = 0;
    num_dofs = subdomain_operator.num_dofs + superdomain_operator.num_dofs - num_interface_dofs;
    interface_is_identity = true;
    const int next = subdomain_operator.num_extended_dofs + superdomain_operator.num_extended_dofs;

    // Norm weighting / inner product weight (tpp:2731-2747)
    norm_weight_hst.assign(next, 1.0);
    for (int i = subdomain_operator.num_dofs; i < subdomain_operator.num_extended_dofs; i++) norm_weight_hst[i] = 0.0;
    norm_weight = device.malloc<DType>(std::max(next, 1));
    norm_weight.copyFrom(norm_weight_hst.data(), next * sizeof(DType));
    num_values = subdomain_operator.num_points + superdomain_operator.num_extended_dofs;
    inner_weight = device.malloc<DType>(std::max(num_values, 1));
    subdomain_operator.Q.multiply(inner_weight, norm_weight);
    inner_weight_hst.resize(num_values);
    inner_weight.copyTo(inner_weight_hst.data(), num_values * sizeof(DType));
    for (auto &w : inner_weight_hst)
        if (w > 0.0) w = 1.0;
    inner_weight.copyFrom(inner_weight_hst.data(), num_values * sizeof(DType));

    // Low-order preconditioner (tpp:2749-3549)
    rstdout("Assembling subdomain low-order preconditioner\n");
    if (use_preconditioner)
    {
        setup_mark("region, Q, weights (1 rank)");
        assemble_low_order_fem();
        setup_mark("low-order FEM assembly");
        amg_fem.coarsening = opt.amg_coarsening;
        amg_fem.fp32 = opt.amg_precision == 1;
        amg_fem.setup(A_fem_hst, cheby_order);
        setup_mark("AMG #2 (hierarchy, upload, collapsed coarse levels)");
    }
}

// A_sub_fem / A_fem for a conforming region (tpp:2913-3472).  Every GLL cell is split in 2 triangles / 6
// tetrahedra; P1 stiffness per simplex; entries of a simplex matrix with |v| <= epsilon are dropped
// (tpp:3025); the per-element matrix is accumulated in loop order, then scattered to the dofs.
template <typename DType>
void Subdomain<DType>::assemble_low_order_fem()
{
    using namespace prfdd_host;
    const int num_verts = (dim == 2) ? 3 : 4;
    const DType weight = (dim == 2) ? 6.0 : 24.0;
    static const int tri[2][3][3] = {{{0, 0, 0}, {1, 0, 0}, {1, 1, 0}}, {{1, 1, 0}, {0, 1, 0}, {0, 0, 0}}};
    static const int tet[6][4][3] = {{{0, 0, 0}, {0, 1, 0}, {1, 0, 0}, {1, 0, 1}}, {{1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {1, 0, 1}}, {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {1, 0, 1}},
                                     {{1, 0, 1}, {1, 1, 0}, {1, 1, 1}, {0, 1, 0}}, {{0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {0, 1, 0}}, {{1, 0, 1}, {1, 1, 1}, {0, 1, 1}, {0, 1, 0}}};
    const int num_low = (dim == 2) ? 2 : 6;
    // D_fem[m][q][v]: gradient of the barycentric basis, identical at every quadrature point (tpp:2833-2843)
    DType D_fem[3][4][4];
    memset(D_fem, 0, sizeof(D_fem));
    for (int q = 0; q < num_verts; q++)
    {
        if (dim == 2)
        {
            D_fem[0][q][0] = -1.0; D_fem[0][q][1] = 1.0; D_fem[0][q][2] = 0.0;
            D_fem[1][q][0] = -1.0; D_fem[1][q][1] = 0.0; D_fem[1][q][2] = 1.0;
        }
        else
        {
            D_fem[0][q][0] = 1.0; D_fem[0][q][3] = -1.0;
            D_fem[1][q][1] = 1.0; D_fem[1][q][3] = -1.0;
            D_fem[2][q][2] = 1.0; D_fem[2][q][3] = -1.0;
        }
    }

    const int n_ext = subdomain_operator.num_extended_dofs;
    std::vector<std::vector<std::pair<int, DType>>> rows(n_ext); // (col, value) in insertion order

    const int nslots = (dim == 2) ? 9 : 27;
    // The elements are independent until their contributions are appended to the rows; the P1 simplex matrices (6 tetrahedra per
    // GLL cell) dominate the set-up time.  Chunks of elements are processed by a few threads, each emitting (row, col, value)
    // triplets in the serial order; the triplets are then appended chunk by chunk, so every row receives its entries in exactly
    // the order of the serial loop and the assembled matrix is bit-identical for any thread count.
    struct Trip
    {
        int row, col;
        DType val;
    };
    const int num_region = (int)subdomain_region.size();
    long long region_points = 0;
    for (auto &el : subdomain_region) region_points += el.num_points;
    const int T = amg::host_threads((int)std::min<long long>(region_points, 1 << 30));
    const int nchunks = T == 1 ? 1 : std::min(num_region, 16 * T);
    std::vector<std::vector<Trip>> emitted(std::max(nchunks, 1));
    std::atomic<int> next_chunk(0);
    auto worker = [&]() {
    std::vector<DType> Ae; // [point][slot], slot = neighbour offset (di,dj,dk) in {-1,0,1}^dim
    std::vector<char> touched;
    for (int chunk = next_chunk++; chunk < nchunks; chunk = next_chunk++)
    {
    std::vector<Trip> &out = emitted[chunk];
    const int e_lo = (int)((long long)num_region * chunk / nchunks), e_hi = (int)((long long)num_region * (chunk + 1) / nchunks);
    for (int e_idx = e_lo; e_idx < e_hi; e_idx++)
    {
        Element<DType> &elem_i = subdomain_region[e_idx];
        const int N_i = elem_i.poly_degree, n_i = N_i + 1;
        const int npts = elem_i.num_points;
        Ae.assign((size_t)npts * nslots, 0.0);
        touched.assign((size_t)npts * nslots, 0);
        auto slot_of = [&](int a, int b) {
            int ia = a % n_i, ja = (a / n_i) % n_i, ka = a / (n_i * n_i);
            int ib = b % n_i, jb = (b / n_i) % n_i, kb = b / (n_i * n_i);
            return (ib - ia + 1) + (jb - ja + 1) * 3 + (dim == 3 ? (kb - ka + 1) * 9 : 0);
        };
        if (N_i > 1)
        {
            const int S_x = N_i, S_y = N_i, S_z = (dim >= 3) ? N_i : 1;
            int loc_sub[4];
            DType x_sub[4], y_sub[4], z_sub[4], H[9], invH[9];
            for (int s_z = 0; s_z < S_z; s_z++)
                for (int s_y = 0; s_y < S_y; s_y++)
                    for (int s_x = 0; s_x < S_x; s_x++)
                        for (int t = 0; t < num_low; t++)
                        {
                            for (int vid = 0; vid < num_verts; vid++)
                            {
                                const int i = (dim == 2) ? tri[t][vid][0] : tet[t][vid][0];
                                const int j = (dim == 2) ? tri[t][vid][1] : tet[t][vid][1];
                                const int k = (dim == 2) ? 0 : tet[t][vid][2];
                                loc_sub[vid] = (dim == 2) ? (s_x + i) + (s_y + j) * n_i : (s_x + i) + (s_y + j) * n_i + (s_z + k) * (n_i * n_i);
                                x_sub[vid] = elem_i.x[loc_sub[vid]];
                                y_sub[vid] = elem_i.y[loc_sub[vid]];
                                z_sub[vid] = (dim >= 3) ? elem_i.z[loc_sub[vid]] : 0.0;
                            }
                            DType det;
                            if (dim == 2)
                            {
                                H[0] = x_sub[1] - x_sub[0]; H[1] = x_sub[2] - x_sub[0];
                                H[2] = y_sub[1] - y_sub[0]; H[3] = y_sub[2] - y_sub[0];
                                det = H[0] * H[3] - H[1] * H[2];
                                invH[0] = (1.0 / det) * H[3]; invH[1] = -(1.0 / det) * H[1];
                                invH[2] = -(1.0 / det) * H[2]; invH[3] = (1.0 / det) * H[0];
                            }
                            else
                            {
                                H[0] = x_sub[0] - x_sub[3]; H[1] = x_sub[1] - x_sub[3]; H[2] = x_sub[2] - x_sub[3];
                                H[3] = y_sub[0] - y_sub[3]; H[4] = y_sub[1] - y_sub[3]; H[5] = y_sub[2] - y_sub[3];
                                H[6] = z_sub[0] - z_sub[3]; H[7] = z_sub[1] - z_sub[3]; H[8] = z_sub[2] - z_sub[3];
                                det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
                                const DType r = 1.0 / det;
                                invH[0] = r * (H[4] * H[8] - H[7] * H[5]); invH[1] = r * (H[2] * H[7] - H[8] * H[1]); invH[2] = r * (H[1] * H[5] - H[4] * H[2]);
                                invH[3] = r * (H[5] * H[6] - H[8] * H[3]); invH[4] = r * (H[0] * H[8] - H[6] * H[2]); invH[5] = r * (H[2] * H[3] - H[5] * H[0]);
                                invH[6] = r * (H[3] * H[7] - H[6] * H[4]); invH[7] = r * (H[1] * H[6] - H[7] * H[0]); invH[8] = r * (H[0] * H[4] - H[3] * H[1]);
                            }
                            DType Gmn[3][3];
                            for (int m = 0; m < dim; m++)
                                for (int nn = 0; nn < dim; nn++)
                                {
                                    DType G_val = 0.0;
                                    for (int k = 0; k < dim; k++) G_val += (det / weight) * invH[m * dim + k] * invH[nn * dim + k];
                                    Gmn[m][nn] = G_val;
                                }
                            DType At[4][4];
                            for (int i = 0; i < num_verts; i++)
                                for (int j = 0; j < num_verts; j++) At[i][j] = 0.0;
                            for (int m = 0; m < dim; m++)
                                for (int nn = 0; nn < dim; nn++)
                                    for (int q = 0; q < num_verts; q++)
                                        for (int i = 0; i < num_verts; i++)
                                        {
                                            const DType dmi = D_fem[m][q][i];
                                            if (dmi == 0.0) continue;
                                            for (int j = 0; j < num_verts; j++)
                                            {
                                                const DType dnj = D_fem[nn][q][j];
                                                if (dnj == 0.0) continue;
                                                At[i][j] = At[i][j] + dmi * (Gmn[m][nn] * dnj);
                                            }
                                        }
                            for (int i = 0; i < num_verts; i++)
                                for (int j = 0; j < num_verts; j++)
                                    if (std::abs(At[i][j]) > epsilon)
                                    {
                                        const size_t s = (size_t)loc_sub[i] * nslots + slot_of(loc_sub[i], loc_sub[j]);
                                        Ae[s] += At[i][j];
                                        touched[s] = 1;
                                    }
                        }
        }
        else
        {
            // N = 1: Q1 SEM element matrix D^T G D (tpp:3040-3124)
            const int nv = npts;
            const std::vector<DType> &d2 = D_hat[num_levels - 1].first;
            std::vector<DType> Dm[3];
            for (int c = 0; c < dim; c++) Dm[c].assign(nv * nv, 0.0);
            if (dim == 2)
            {
                for (int k = 0; k < 2; k++) for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) Dm[0][(i + k * 2) * 4 + (j + k * 2)] = d2[i * 2 + j];
                for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) for (int k = 0; k < 2; k++) Dm[1][(i * 2 + k) * 4 + (j * 2 + k)] = d2[i * 2 + j];
            }
            else
            {
                for (int p = 0; p < 2; p++) for (int q = 0; q < 2; q++) for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++)
                {
                    Dm[0][(i + (p * 2 + q) * 2) * 8 + (j + (p * 2 + q) * 2)] = d2[i * 2 + j];
                    Dm[1][(i * 8 + j) * 2 + ((p + p * 8) * (2 * 2) + (q + q * 8))] = d2[i * 2 + j];
                    Dm[2][(i * 8 + j) * (2 * 2) + (p + q * 2) * (1 + 8)] = d2[i * 2 + j];
                }
            }
            static const int gi2[2][2] = {{0, 2}, {2, 1}};
            static const int gi3[3][3] = {{0, 3, 4}, {3, 1, 5}, {4, 5, 2}};
            for (int i = 0; i < nv; i++)
                for (int j = 0; j < nv; j++)
                {
                    DType val = 0.0;
                    for (int k = 0; k < nv; k++)
                        for (int a = 0; a < dim; a++)
                        {
                            DType gd = 0.0; // (G D)_a [k][j] = sum_b G_ab[k] D_b[k][j]
                            for (int b = 0; b < dim; b++) gd += elem_i.geom_fact[dim == 2 ? gi2[a][b] : gi3[a][b]][k] * Dm[b][k * nv + j];
                            val += Dm[a][k * nv + i] * gd;
                        }
                    if (std::abs(val) > epsilon)
                    {
                        const size_t s = (size_t)i * nslots + slot_of(i, j);
                        Ae[s] += val;
                        touched[s] = 1;
                    }
                }
        }
        // does any edge / face neighbour in the region have a lower degree? (hanging nodes on this element)
        bool conforming = true;
        for (auto &c : elem_i.edge_conn)
            for (int j : c) conforming = conforming && !(subdomain_region[j].poly_degree < N_i);
        for (auto &c : elem_i.face_conn)
            for (int j : c) conforming = conforming && !(subdomain_region[j].poly_degree < N_i);
        auto slot_target = [&](int a, int s) {
            const int ia = a % n_i, ja = (a / n_i) % n_i, ka = a / (n_i * n_i);
            const int ib = ia + (s % 3) - 1, jb = ja + ((s / 3) % 3) - 1, kb = ka + (dim == 3 ? (s / 9) - 1 : 0);
            return ib + jb * n_i + kb * n_i * n_i;
        };
        if (conforming)
        {
            // J_e is a selection (tpp:3287-3297): scatter A_e to the dofs (tpp:3385-3403)
            for (int a = 0; a < npts; a++)
            {
                const long long ra = elem_i.dof_num[a];
                if (ra <= 0) continue;
                for (int s = 0; s < nslots; s++)
                {
                    if (!touched[(size_t)a * nslots + s]) continue;
                    const DType val = Ae[(size_t)a * nslots + s];
                    if (!(std::abs(val) > epsilon)) continue;
                    const long long cb = elem_i.dof_num[slot_target(a, s)];
                    if (cb <= 0) continue;
                    out.push_back(Trip{(int)(ra - 1), (int)(cb - 1), val});
                }
            }
            continue;
        }
        // hanging nodes: columns of J_e = own points with glo_num > 0, then the coarse neighbours' edge / face interiors
        // (tpp:3130-3355); A_sub_fem += J_e^T A_e J_e
        {
            using namespace prfdd_multi;
            const int num_edges = (dim == 2) ? 4 : 12, num_faces = (dim == 2) ? 0 : 6;
            int rank = 1;
            std::vector<std::pair<int, long long>> vert(npts);
            for (int v = 0; v < npts; v++) vert[v] = {elem_i.glo_num[v] > 0 ? rank++ : 0, elem_i.dof_num[v]};
            std::vector<std::vector<int>> edge_pts(num_edges), face_pts(num_faces);
            std::vector<std::vector<std::pair<int, long long>>> edge_cols(num_edges), face_cols(num_faces);
            for (int q = 0; q < num_edges; q++)
            {
                const int e_j = min_degree_edge_neighbor(subdomain_region, elem_i, q);
                if (e_j < 0) continue;
                const Element<DType> &ej = subdomain_region[e_j];
                const int n_j = ej.poly_degree + 1;
                auto m = matching(elem_i, ej, 0, q);
                edge_pts[q] = m.first;
                edge_cols[q].resize(n_j);
                edge_cols[q][0] = vert[m.first[0]];
                edge_cols[q][n_j - 1] = vert[m.first[n_i - 1]];
                for (int k = 1; k < n_j - 1; k++) edge_cols[q][k] = {rank++, ej.dof_num[m.second[k]]};
            }
            for (int q = 0; q < num_faces; q++)
                for (int e_j : elem_i.face_conn[q])
                {
                    const Element<DType> &ej = subdomain_region[e_j];
                    const int n_j = ej.poly_degree + 1;
                    if (N_i <= ej.poly_degree) continue;
                    auto m = matching(elem_i, ej, 1, q);
                    face_pts[q] = m.first;
                    auto &lst = face_cols[q];
                    lst.assign(n_j * n_j, {0, 0});
                    lst[0] = vert[m.first[0]];
                    lst[n_j - 1] = vert[m.first[n_i - 1]];
                    lst[(n_j - 1) * n_j] = vert[m.first[(n_i - 1) * n_i]];
                    lst[n_j * n_j - 1] = vert[m.first[n_i * n_i - 1]];
                    const int *fe = FACE_EDGES[q];
                    for (int k = 1; k < n_j - 1; k++)
                    {
                        lst[k] = edge_cols[fe[0]][k];
                        lst[k + (n_j - 1) * n_j] = edge_cols[fe[1]][k];
                        lst[k * n_j] = edge_cols[fe[2]][k];
                        lst[(n_j - 1) + k * n_j] = edge_cols[fe[3]][k];
                    }
                    for (int b = 1; b < n_j - 1; b++)
                        for (int a = 1; a < n_j - 1; a++) lst[a + b * n_j] = {rank++, ej.dof_num[m.second[a + b * n_j]]};
                }
            const int ncols = rank - 1;
            std::vector<std::vector<std::pair<int, DType>>> Je(npts);
            std::vector<long long> dcol(ncols, 0);
            for (int v = 0; v < npts; v++)
                if (vert[v].first > 0) { Je[v].push_back({vert[v].first - 1, 1.0}); dcol[vert[v].first - 1] = vert[v].second; }
            for (int q = 0; q < num_edges; q++)
            {
                if (edge_cols[q].empty()) continue;
                const int n_j = (int)edge_cols[q].size();
                const std::vector<DType> &Jf = J_cf_fem.at(std::pair<int, int>(n_j - 1, N_i));
                for (auto &pr : edge_cols[q]) dcol[pr.first - 1] = pr.second;
                for (int i = 1; i < n_i - 1; i++)
                    for (int j = 0; j < n_j; j++)
                        if (std::abs(Jf[i * n_j + j]) > epsilon) Je[edge_pts[q][i]].push_back({edge_cols[q][j].first - 1, Jf[i * n_j + j]});
            }
            for (int q = 0; q < num_faces; q++)
            {
                if (face_cols[q].empty()) continue;
                const int n_j = (int)std::lround(std::sqrt((double)face_cols[q].size()));
                const std::vector<DType> &Jf = J_cf_fem.at(std::pair<int, int>(n_j - 1, N_i));
                for (auto &pr : face_cols[q]) dcol[pr.first - 1] = pr.second;
                for (int j = 1; j < n_i - 1; j++)
                    for (int i = 1; i < n_i - 1; i++)
                        for (int qq = 0; qq < n_j; qq++)
                            for (int pp = 0; pp < n_j; pp++)
                            {
                                const DType val = Jf[i * n_j + pp] * Jf[j * n_j + qq];
                                if (std::abs(val) > epsilon) Je[face_pts[q][i + j * n_i]].push_back({face_cols[q][pp + qq * n_j].first - 1, val});
                            }
            }
            std::vector<DType> acc((size_t)ncols * ncols, 0.0);
            for (int a = 0; a < npts; a++)
                for (int s = 0; s < nslots; s++)
                {
                    if (!touched[(size_t)a * nslots + s]) continue;
                    const DType val = Ae[(size_t)a * nslots + s];
                    const int b = slot_target(a, s);
                    for (auto &ca : Je[a])
                        for (auto &cb : Je[b]) acc[(size_t)ca.first * ncols + cb.first] += ca.second * val * cb.second;
                }
            for (int ca = 0; ca < ncols; ca++)
            {
                if (dcol[ca] <= 0) continue;
                for (int cb = 0; cb < ncols; cb++)
                {
                    const DType val = acc[(size_t)ca * ncols + cb];
                    if (std::abs(val) > epsilon && dcol[cb] > 0) out.push_back(Trip{(int)(dcol[ca] - 1), (int)(dcol[cb] - 1), val});
                }
            }
        }
    }
    }
    };
    if (T == 1) worker();
    else
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < T; t++) pool.emplace_back(worker);
        for (auto &th : pool) th.join();
    }
    for (auto &chunk : emitted)
    {
        for (const Trip &t : chunk) rows[t.row].push_back({t.col, t.val});
        std::vector<Trip>().swap(chunk);
    }

    // A_sub_fem rows -> composite A_fem through the interface numbering (tpp:3414-3472)
    const int ne = subdomain_operator.num_extended_dofs;
    std::vector<int> m(ne + superdomain_operator.num_extended_dofs);
    if (interface_is_identity)
        for (size_t i = 0; i < m.size(); i++) m[i] = (int)i;
    else
        for (size_t i = 0; i < m.size(); i++) m[i] = Q_int.col_hst[Q_int.ptr_hst[i]];
    std::vector<std::tuple<int, int, double>> coo;
    for (int i = 0; i < subdomain_operator.num_dofs; i++)
    {
        auto &row = rows[i];
        std::stable_sort(row.begin(), row.end(), [](const std::pair<int, DType> &a, const std::pair<int, DType> &b) { return a.first < b.first; });
        for (size_t k = 0; k < row.size(); k++)
        {
            if (k > 0 && row[k].first == row[k - 1].first)
                std::get<2>(coo.back()) += row[k].second;
            else
                coo.emplace_back(m[i], m[row[k].first], row[k].second);
        }
        std::vector<std::pair<int, DType>>().swap(row);
    }
    if (superdomain_operator.num_dofs > 0)
    {
        auto &As = superdomain_operator.A;
        for (int i = num_interface_dofs; i < superdomain_operator.num_dofs; i++)
            for (int k = As.ptr_hst[i]; k < As.ptr_hst[i + 1]; k++) coo.emplace_back(m[ne + i], m[ne + As.col_hst[k]], As.val_hst[k]);
    }
    A_fem_hst = prfdd_multi::csr_from_coo(num_dofs, num_dofs, coo);
}

template <typename DType>
void Subdomain<DType>::allocate_solver()
{
    using namespace prfdd_host;
    // Solver (tpp:3857-3873)
    const int nvl = std::max(num_values, 1);
    f = device.malloc<DType>(nvl); u_k = device.malloc<DType>(nvl); r_k = device.malloc<DType>(nvl); r_kp1 = device.malloc<DType>(nvl);
    q_k = device.malloc<DType>(nvl); z_k = device.malloc<DType>(nvl); p_k = device.malloc<DType>(nvl);
    V.resize(num_vectors + 1);
    for (auto &v : V) v = device.malloc<DType>(nvl);
    Z.resize(num_vectors);
    for (auto &z : Z) z = device.malloc<DType>(nvl);
    const int next = std::max(subdomain_operator.num_extended_dofs + superdomain_operator.num_extended_dofs, 1);
    aV.resize(num_vectors + 1);
    for (auto &v : aV) v = device.malloc<DType>(next);
    aq = device.malloc<DType>(next);
    int total_level_points = levels.back().offset + levels.back().num_points;
    int wsize = std::max({nvl, next, total_level_points, 1});
    work_dev.resize(3);
    for (auto &w : work_dev) w = device.malloc<DType>(wsize);
    kstate = device.malloc<char>(sizeof(prfdd_krylov_state));
    prfdd_krylov_state zero;
    memset(&zero, 0, sizeof(zero));
    zero.one = 1.0;
    kstate.copyFrom(&zero, sizeof(zero));
}

// ---------------------------------------------------------------------------------------------
// run time
// ---------------------------------------------------------------------------------------------
template <typename DType>
void Subdomain<DType>::stiffness_matrix(const memory &Au, const memory &u)
{
    // subdomain.tpp:3942-3967
    const int npt = subdomain_operator.num_points, ns = superdomain_operator.num_extended_dofs;
    if (ns > 0)
    {
        memory u_sup = u.slice(npt, ns), Au_sup = Au.slice(npt, ns);
        superdomain_operator.A.multiply(Au_sup, u_sup);
    }
    const double *g[6];
    for (int c = 0; c < 6; c++) g[c] = dp(subdomain_operator.geom_fact[c]);
    dev::check_rc(prfdd_stiffness_matrix_region_hd(dp(Au), dp(u), g, (int)subdomain_operator.bucket_n.size(), subdomain_operator.bucket_first_point.data(),
                                                   subdomain_operator.bucket_num_elements.data(), subdomain_operator.bucket_n.data(), subdomain_operator.bucket_D.data(),
                                                   subdomain_operator.bucket_D_hst.data(), prfdd_host::dim, st()),
                  "Subdomain::stiffness_matrix");
}

template <typename DType>
void Subdomain<DType>::direct_stiffness_summation(const memory &QQtu, const memory &u)
{
    // subdomain.tpp:3969-3985
    const int npt = subdomain_operator.num_points, ne = subdomain_operator.num_extended_dofs, ns = superdomain_operator.num_extended_dofs;
    memory u_sub = u.slice(0, npt), out_sub = QQtu.slice(0, npt);
    subdomain_operator.Qt.multiply(work_dev[0], u_sub);
    if (ns > 0) u.slice(npt, ns).copyTo(work_dev[0].slice(ne, ns), ns * sizeof(DType));
    if (interface_is_identity)
    {
        // QQt_int = identity on the dofs, zero on the extended dofs (tpp:2677-2681)
        if (ne > subdomain_operator.num_dofs) math.set_to_value(work_dev[0], 0.0, ne - subdomain_operator.num_dofs, subdomain_operator.num_dofs);
        subdomain_operator.Q.multiply(out_sub, work_dev[0]);
    }
    else
    {
        QQt_int.multiply(work_dev[1], work_dev[0]);
        subdomain_operator.Q.multiply(out_sub, work_dev[1]);
        if (ns > 0) QQtu.slice(npt, ns).copyFrom(work_dev[1].slice(ne, ns), ns * sizeof(DType));
    }
}

template <typename DType>
void Subdomain<DType>::low_order_preconditioner(const memory &z, const memory &r, const memory *r_assembled)
{
    using namespace prfdd_host;
    // subdomain.tpp:3987-4159: z = Q Q_int Vcycle(A_fem) Qt_int Qt r
    // r_assembled: [Qt r_sub | r_sup] when the caller already holds it (the Arnoldi loop keeps the assembled copy of every
    // basis vector for its inner products), which saves the assembly pass here
    const int npt = subdomain_operator.num_points, ne = subdomain_operator.num_extended_dofs, ns = superdomain_operator.num_extended_dofs;
    memory r_sub = r.slice(0, npt), z_sub = z.slice(0, npt);
    amg::Level &L0 = amg_fem.levels[0];
    if (interface_is_identity)
    {
        // Qt_int / Q_int are identities: assemble straight into the V-cycle's right-hand side, scatter straight out of its solution
        if (amg_fem.fp32)
        {
            // FP32 cycle under the FP64 Krylov method: cast in, cycle, cast out (copy_from/to_domain_data with EType != DType)
            const memory *src = r_assembled;
            if (!src)
            {
                subdomain_operator.Qt.multiply(work_dev[0], r_sub);
                src = &work_dev[0];
            }
            dev::check_rc(prfdd_cast_f64_to_f32(L0.f.as<float>(), dp(*src), L0.n, st()), "cast to FP32");
            timer.start("subdomain.preconditioner.down_leg_gpu");
            amg_fem.vcycle(num_vcycles);
            timer.stop("subdomain.preconditioner.down_leg_gpu");
            dev::check_rc(prfdd_cast_f32_to_f64(dp(work_dev[0]), L0.u.as<float>(), L0.n, st()), "cast to FP64");
            subdomain_operator.Q.multiply(z_sub, work_dev[0]);
            return;
        }
        memory f_own = L0.f;
        if (r_assembled)
            L0.f = r_assembled->slice(0, L0.n); // the cycle reads its right-hand side in place
        else
        {
            timer.start("subdomain.preconditioner.assemble_subdomain");
            subdomain_operator.Qt.multiply(L0.f, r_sub);
            timer.stop("subdomain.preconditioner.assemble_subdomain");
        }
        timer.start("subdomain.preconditioner.down_leg_gpu");
        amg_fem.vcycle(num_vcycles);
        timer.stop("subdomain.preconditioner.down_leg_gpu");
        L0.f = f_own;
        timer.start("subdomain.preconditioner.unassemble_subdomain");
        subdomain_operator.Q.multiply(z_sub, L0.u);
        timer.stop("subdomain.preconditioner.unassemble_subdomain");
        return;
    }
    if (!r_assembled)
    {
        timer.start("subdomain.preconditioner.assemble_subdomain");
        subdomain_operator.Qt.multiply(work_dev[0], r_sub);
        timer.stop("subdomain.preconditioner.assemble_subdomain");
        if (ns > 0) work_dev[0].slice(ne, ns).copyFrom(r.slice(npt, ns), ns * sizeof(DType));
    }
    // interface assembly fused with the head of level 0's zero-guess smoothing: f = Qt_int a, r = ds f, t0 = ds (c r) in one pass
    timer.start("subdomain.preconditioner.assemble_composite");
    const bool head = !amg_fem.fp32 && amg_fem.num_levels() > 1 && Qt_int.num_rows == L0.n && Qt_int.num_nnz > 0 && !Qt_int.unit_values;
    if (amg_fem.fp32)
    {
        Qt_int.multiply(work_dev[1], r_assembled ? *r_assembled : work_dev[0]);
        dev::check_rc(prfdd_cast_f64_to_f32(L0.f.as<float>(), dp(work_dev[1]), L0.n, st()), "cast to FP32");
    }
    else if (head)
        dev::check_rc(prfdd_csrm_restrict_cheby_residual(dp(L0.f), dp(L0.r), dp(L0.t0), &Qt_int.desc, dp(r_assembled ? *r_assembled : work_dev[0]), dp(L0.ds),
                                                         L0.coefs[amg_fem.cheby_order - 1], st()), "Qt_int + smoothing head");
    else
        Qt_int.multiply(L0.f, r_assembled ? *r_assembled : work_dev[0]);
    timer.stop("subdomain.preconditioner.assemble_composite");
    timer.start("subdomain.preconditioner.down_leg_gpu");
    amg_fem.vcycle(num_vcycles, head);
    timer.stop("subdomain.preconditioner.down_leg_gpu");
    timer.start("subdomain.preconditioner.unassemble_composite");
    if (amg_fem.fp32)
    {
        dev::check_rc(prfdd_cast_f32_to_f64(dp(work_dev[1]), L0.u.as<float>(), L0.n, st()), "cast to FP64");
        Q_int.multiply(work_dev[0], work_dev[1]);
    }
    else
        Q_int.multiply(work_dev[0], L0.u);
    timer.stop("subdomain.preconditioner.unassemble_composite");
    timer.start("subdomain.preconditioner.unassemble_subdomain");
    subdomain_operator.Q.multiply(z_sub, work_dev[0]);
    timer.stop("subdomain.preconditioner.unassemble_subdomain");
    if (ns > 0) z.slice(npt, ns).copyFrom(work_dev[0].slice(ne, ns), ns * sizeof(DType));
}

// dst[0:ne] = norm_weight .* (Qt src_sub) ; dst[ne:ne+ns] = src_sup       (tpp:4285-4286, 4500-4501)
template <typename DType>
void Subdomain<DType>::assemble_weighted(const memory &dst, const memory &src)
{
    const int npt = subdomain_operator.num_points, ne = subdomain_operator.num_extended_dofs, ns = superdomain_operator.num_extended_dofs;
    subdomain_operator.Qt.multiply_weight(dst, src.slice(0, npt), norm_weight);
    if (ns > 0) src.slice(npt, ns).copyTo(dst.slice(ne, ns), ns * sizeof(DType));
}

// dst[0:ne] = Qt src_sub ; dst[ne:ne+ns] = src_sup: the assembled copy without the 0/1 norm weight, which the inner products
// apply themselves (w in {0,1}, so sum w (w a)(w b) and sum w a b are the same number); it is also exactly the vector the
// low-order preconditioner assembles from its input (tpp:3994-3999)
template <typename DType>
void Subdomain<DType>::assemble(const memory &dst, const memory &src)
{
    const int npt = subdomain_operator.num_points, ne = subdomain_operator.num_extended_dofs, ns = superdomain_operator.num_extended_dofs;
    subdomain_operator.Qt.multiply(dst, src.slice(0, npt));
    if (ns > 0) src.slice(npt, ns).copyTo(dst.slice(ne, ns), ns * sizeof(DType));
}

template <typename DType>
void Subdomain<DType>::residual_norm_dev(const memory &r, double *out)
{
    // subdomain.tpp:4491-4515, without the D2H + host sum: out[0] = sum w (w Qt r)^2 on the device
    const int next = subdomain_operator.num_extended_dofs + superdomain_operator.num_extended_dofs;
    assemble_weighted(aq, r);
    dev::check_rc(prfdd_weighted_inner_product(ws, out, dp(aq), dp(aq), dp(norm_weight), next, st()), "Subdomain::residual_norm");
}

template <typename DType>
void Subdomain<DType>::tree_operator(const memory &Tu, const memory &u)
{
    using namespace prfdd_host;
    // subdomain.tpp:4566-4646
    timer.start("subdomain.tree_construction.gpu_to_gpu");
    if (num_procs == 1)
    {
        // own elements at degree N are the whole region; the ladder restrictions (tpp:4576-4609) would feed only
        // other ranks' regions and the empty superdomain, so they are not launched
        dev::check_rc(prfdd_copy_from_domain_data(dp(Tu), dp(u), own_points, st()), "copy_from_domain_data");
        timer.stop("subdomain.tree_construction.gpu_to_gpu");
        return;
    }
    timer.stop("subdomain.tree_construction.gpu_to_gpu");
    tree_operator_multi(Tu, u);
}

template <typename DType>
void Subdomain<DType>::gmres_body(const memory &u_l, const memory &f_l)
{
    using namespace prfdd_host;
    // subdomain.tpp:4309-4489; all scalars in the device-side prfdd_krylov_state
    const int nvl = num_values;
    const int next = subdomain_operator.num_extended_dofs + superdomain_operator.num_extended_dofs;
    prfdd_krylov_state *K = ks();
    double *red = ks_ptr(offsetof(prfdd_krylov_state, red));
    double *hcol = ks_ptr(offsetof(prfdd_krylov_state, hcol));
    double *inv_gamma0 = ks_ptr(offsetof(prfdd_krylov_state, inv_gamma0));
    double *inv_alpha = ks_ptr(offsetof(prfdd_krylov_state, inv_alpha));
    double *ycoef = ks_ptr(offsetof(prfdd_krylov_state, y));

    tree_operator(f, f_l);
    dev::check_rc(prfdd_krylov_reset(K, st()), "krylov_reset");

    timer.start("subdomain.vector_operations");
    dev::check_rc(prfdd_initialize_arrays(dp(u_k), dp(r_k), dp(f), nvl, st()), "initialize_arrays");
    timer.stop("subdomain.vector_operations");

    // aV[i] holds the assembled copy [Qt V_i,sub | V_i,sup] of every basis vector WITHOUT the norm weight (assemble()); the
    // weight enters the reductions.  One Arnoldi column is then: V-cycle on aV[j] directly, operator, ONE assembly of the new
    // column, one fused multi-dot, one fused orthogonalise+norm on the assembled side, the device-side Hessenberg update, and
    // one launch that forms V[j+1] and aV[j+1] (the reference: 2(j+1)+2 assemblies, j+3 host-synchronising reductions, j+3 axpys)
    int iter = 0;
    bool first = true;
    while (iter < max_iterations)
    {
        if (!first)
        {
            stiffness_matrix(r_k, u_k);
            math.vector_vector_addition(r_k, 1.0, f, -1.0, r_k, nvl);
        }
        timer.start("subdomain.residual_norm");
        assemble(aq, r_k);
        dev::check_rc(prfdd_weighted_inner_product(ws, red, dp(aq), dp(aq), dp(norm_weight), next, st()), "Subdomain::residual_norm");
        dev::check_rc(prfdd_gmres_begin_cycle(K, first ? 1 : 0, st()), "gmres_begin_cycle");
        timer.stop("subdomain.residual_norm");
        first = false;

        // V0 = r / gamma0 and its assembled copy
        dev::check_rc(prfdd_arnoldi_next(dp(V[0]), dp(r_k), nullptr, hcol, 0, inv_gamma0, nvl, dp(aV[0]), dp(aq), next, st()), "V0");

        int j;
        for (j = 0; j < num_vectors; j++)
        {
            iter++;
            timer.start("subdomain.preconditioner");
            if (use_preconditioner)
                low_order_preconditioner(Z[j], V[j], &aV[j]);
            else
                direct_stiffness_summation(Z[j], V[j]);
            timer.stop("subdomain.preconditioner");

            timer.start("subdomain.operator_application");
            stiffness_matrix(q_k, Z[j]);
            timer.stop("subdomain.operator_application");

            // Gram-Schmidt (tpp:4389-4401): H[i][j] = <Qt q, Qt V_i>_w for i <= j in one pass
            timer.start("subdomain.inner_products");
            assemble(aq, q_k);
            std::vector<const double *> ap(j + 1), vp(j + 1);
            for (int i = 0; i <= j; i++) { ap[i] = dp(aV[i]); vp[i] = dp(V[i]); }
            dev::check_rc(prfdd_multi_inner_product(ws, hcol, dp(aq), ap.data(), dp(norm_weight), j + 1, next, st()), "multi_inner_product");
            timer.stop("subdomain.inner_products");

            // orthogonalise the assembled copy and take its norm (tpp:4396-4412) -- Qt is linear, no second assembly
            timer.start("subdomain.residual_norm");
            dev::check_rc(prfdd_orthogonalize_norm(ws, red, dp(aq), ap.data(), hcol, dp(norm_weight), j + 1, next, st()), "orthogonalize_norm");
            dev::check_rc(prfdd_gmres_column(K, j, iter, max_iterations, tolerance, 0, st()), "gmres_column");
            timer.stop("subdomain.residual_norm");

            if (iter >= max_iterations) { j++; break; } // statically known: the reference breaks here too (tpp:4449-4453)
            timer.start("subdomain.vector_operations");
            if (j + 1 < num_vectors + 1)
                dev::check_rc(prfdd_arnoldi_next(dp(V[j + 1]), dp(q_k), vp.data(), hcol, j + 1, inv_alpha, nvl, dp(aV[j + 1]), dp(aq), next, st()), "V_j+1");
            timer.stop("subdomain.vector_operations");
        }
        dev::check_rc(prfdd_gmres_end_cycle(K, num_vectors, st()), "gmres_end_cycle");
        // Sum Arnoldi vectors (tpp:4472-4478); coefficients beyond the last used column are exactly 0 and skipped
        {
            int cnt = std::min(j, num_vectors);
            std::vector<const double *> zp(cnt);
            for (int i = 0; i < cnt; i++) zp[i] = dp(Z[i]);
            dev::check_rc(prfdd_multi_axpy_dev(dp(u_k), zp.data(), ycoef, 1, 1.0, cnt, nvl, st()), "sum Arnoldi");
        }
    }
    dev::check_rc(prfdd_copy_to_domain_data(dp(u_l), dp(u_k), own_points, st()), "copy_to_domain_data"); // tpp:4485
}

template <typename DType>
void Subdomain<DType>::fcg_body(const memory &u_l, const memory &f_l)
{
    using namespace prfdd_host;
    // subdomain.tpp:4161-4268; scalars on the device
    const int nvl = num_values;
    prfdd_krylov_state *K = ks();
    double *red = ks_ptr(offsetof(prfdd_krylov_state, red));
    double *alpha = ks_ptr(offsetof(prfdd_krylov_state, alpha_cg));
    double *beta = ks_ptr(offsetof(prfdd_krylov_state, beta_cg));
    double *one = ks_ptr(offsetof(prfdd_krylov_state, one));
    int *stopped = reinterpret_cast<int *>(reinterpret_cast<char *>(kstate.ptr()) + offsetof(prfdd_krylov_state, stopped));

    tree_operator(r_k, f_l);
    dev::check_rc(prfdd_krylov_reset(K, st()), "krylov_reset");
    math.set_to_value(u_k, 0.0, nvl);
    residual_norm_dev(r_k, red + 2); // r_0_norm (only used by a relative test, which the reference never selects)

    auto precond = [&](const memory &z, const memory &r) {
        if (use_preconditioner) low_order_preconditioner(z, r);
        else direct_stiffness_summation(z, r);
    };
    precond(z_k, r_k);
    p_k.copyFrom(z_k, nvl * sizeof(DType));

    int iter = 0;
    while (iter < max_iterations)
    {
        stiffness_matrix(q_k, p_k);
        dev::check_rc(prfdd_weighted_projection_inner_products(ws, red, dp(z_k), dp(r_k), dp(p_k), dp(q_k), dp(inner_weight), nvl, st()), "projection_inner_products");
        dev::check_rc(prfdd_fcg_alpha(K, st()), "fcg_alpha");
        dev::check_rc(prfdd_solution_and_residual_update_dev(dp(u_k), dp(r_kp1), dp(r_k), dp(p_k), dp(q_k), alpha, one, nvl, st()), "solution_and_residual_update");
        residual_norm_dev(r_kp1, red + 2);
        iter++;
        dev::check_rc(prfdd_fcg_check(K, iter, max_iterations, tolerance, 0, st()), "fcg_check");
        if (iter == max_iterations) break;
        precond(z_k, r_kp1);
        dev::check_rc(prfdd_search_update_inner_product(ws, red + 3, dp(r_k), dp(r_kp1), dp(z_k), dp(inner_weight), nvl, st()), "search_update_inner_product");
        dev::check_rc(prfdd_fcg_beta(K, st()), "fcg_beta");
        dev::check_rc(prfdd_residual_and_search_update_gated(dp(p_k), dp(r_k), dp(z_k), dp(r_kp1), beta, stopped, nvl, st()), "residual_and_search_update");
    }
    dev::check_rc(prfdd_copy_to_domain_data(dp(u_l), dp(u_k), own_points, st()), "copy_to_domain_data"); // tpp:4266
}

// Runs `body` directly, or -- with use_cuda_graph -- captures it once per (input, output) pair and replays it.
template <typename DType>
template <class Body>
void Subdomain<DType>::run_captured(int type, const memory &u_l, const memory &f_l, Body body)
{
    using namespace prfdd_host;
    {
        // called while the stream is being captured (the outer solve's graph): the body's launches go straight into that graph
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        dev::check(cudaStreamIsCapturing(st(), &cs), "cudaStreamIsCapturing");
        if (cs == cudaStreamCaptureStatusActive)
        {
            body();
            return;
        }
    }
    if (!opt.use_cuda_graph || timer.enabled)
    {
        long long before = prfdd_launch_count();
        const double bytes_before = prfdd_algorithmic_bytes();
        body();
        launches_per_apply = prfdd_launch_count() - before;
        bytes_per_apply = prfdd_algorithmic_bytes() - bytes_before;
        return;
    }
    GraphKey key{f_l.ptr(), u_l.ptr(), type};
    auto it = graphs.find(key);
    if (it == graphs.end())
    {
        // make sure lazily created state (constant-bank copies of D, function attributes) exists before capturing
        body();
        dev::check(cudaStreamSynchronize(st()), "graph warm-up");
        cudaGraph_t graph;
        long long before = prfdd_launch_count();
        const double bytes_before = prfdd_algorithmic_bytes();
        dev::check(cudaStreamBeginCapture(st(), cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
        body();
        dev::check(cudaStreamEndCapture(st(), &graph), "cudaStreamEndCapture");
        launches_per_apply = prfdd_launch_count() - before;
        bytes_per_apply = prfdd_algorithmic_bytes() - bytes_before;
        prfdd_launch_count_add(-launches_per_apply); // the captured run launched nothing; the warm-up run above did the work
        prfdd_algorithmic_bytes_add(-bytes_per_apply);
        cudaGraphExec_t exec;
        dev::check(cudaGraphInstantiate(&exec, graph, 0), "cudaGraphInstantiate");
        cudaGraphDestroy(graph);
        it = graphs.emplace(key, exec).first;
        // the warm-up run already produced the result for this call
        return;
    }
    dev::check(cudaGraphLaunch(it->second, st()), "cudaGraphLaunch");
    // the kernels inside the graph are this library's launches too
    prfdd_launch_count_add(launches_per_apply);
    prfdd_algorithmic_bytes_add(bytes_per_apply);
}

template <typename DType>
void Subdomain<DType>::generalized_minimum_residual(memory &u_l, memory &f_l, bool, bool)
{
    run_captured(1, u_l, f_l, [&]() { gmres_body(u_l, f_l); });
}

template <typename DType>
void Subdomain<DType>::flexible_conjugate_gradient(memory &u_l, memory &f_l, bool, bool)
{
    run_captured(0, u_l, f_l, [&]() { fcg_body(u_l, f_l); });
}

// ---------------------------------------------------------------------------------------------
// C ABI support
// ---------------------------------------------------------------------------------------------
template <typename DType>
long long Subdomain<DType>::query(int what)
{
    switch (what)
    {
    case PRFDD_Q_SUB_NUM_POINTS: return subdomain_operator.num_points;
    case PRFDD_Q_SUB_NUM_DOFS: return subdomain_operator.num_dofs;
    case PRFDD_Q_SUB_NUM_EXTENDED_DOFS: return subdomain_operator.num_extended_dofs;
    case PRFDD_Q_SUP_NUM_DOFS: return superdomain_operator.num_dofs;
    case PRFDD_Q_SUP_NUM_EXTENDED_DOFS: return superdomain_operator.num_extended_dofs;
    case PRFDD_Q_NUM_VALUES: return num_values;
    case PRFDD_Q_NUM_DOFS: return num_dofs;
    case PRFDD_Q_AMG_NUM_LEVELS: return amg_fem.num_levels();
    case PRFDD_Q_GPU_LAUNCHES_PER_PRECOND: return launches_per_apply;
    case PRFDD_Q_INNER_ITERATIONS:
    {
        if (!kstate.is_initialized()) return 0;
        prfdd_krylov_state h;
        kstate.copyTo(&h, sizeof(h));
        return h.iterations;
    }
    default: return -1;
    }
}

template <typename DType>
int Subdomain<DType>::output(const std::string &output_name, const std::vector<std::pair<std::string, const memory *>> &fields)
{
    // element-major points of the region in the order of `elements` (tpp:4666-4668); a field lives in the region's point numbering:
    // work_hst[1][elem.offset + v] = field[elem.loc_num[v]] (tpp:4779-4781)
    long long num_points = 0;
    for (auto &elem : elements) num_points += elem.num_points;
    std::vector<int> n_of((size_t)elements.size());
    std::vector<double> x((size_t)num_points), y((size_t)num_points), z((size_t)num_points), degree((size_t)num_points);
    std::vector<long long> off(elements.size() + 1, 0);
    for (size_t e = 0; e < elements.size(); e++)
    {
        const auto &elem = elements[e];
        n_of[e] = elem.poly_degree + 1;
        off[e + 1] = off[e] + elem.num_points;
        for (int v = 0; v < elem.num_points; v++)
        {
            x[off[e] + v] = elem.x[v];
            y[off[e] + v] = elem.y.empty() ? 0.0 : elem.y[v];
            z[off[e] + v] = elem.z.empty() ? 0.0 : elem.z[v];
            degree[off[e] + v] = elem.poly_degree;
        }
    }
    std::vector<std::vector<double>> data(fields.size());
    std::vector<const char *> names{"degree"};
    std::vector<const double *> ptrs{degree.data()};
    std::vector<double> raw;
    for (size_t k = 0; k < fields.size(); k++)
    {
        const long long have = (long long)(fields[k].second->size() / sizeof(DType));
        raw.assign((size_t)have, 0.0);
        fields[k].second->copyTo(raw.data(), (size_t)have * sizeof(DType));
        data[k].assign((size_t)num_points, 0.0);
        for (size_t e = 0; e < elements.size(); e++)
            for (int v = 0; v < elements[e].num_points; v++)
            {
                const long long src = elements[e].loc_num[v];
                if (src >= 0 && src < have) data[k][off[e] + v] = raw[src];
            }
        names.push_back(fields[k].first.c_str());
        ptrs.push_back(data[k].data());
    }
    char path[1024];
    snprintf(path, sizeof(path), "%s_%d.vtk", output_name.c_str(), prfdd_host::proc_id);
    return prfdd_write_vtk_mixed(path, prfdd_host::dim, (int)elements.size(), n_of.data(), x.data(), y.data(), z.data(), (int)names.size(), names.data(), ptrs.data());
}

template <typename DType>
long long Subdomain<DType>::get_array(int what, void *dst, long long cap)
{
    auto put = [&](const void *src, size_t elem, long long count) -> long long {
        if ((long long)(elem * count) > cap) return -(long long)(elem * count);
        memcpy(dst, src, elem * count);
        return count;
    };
    switch (what)
    {
    case PRFDD_A_SUB_Q_PTR: return put(subdomain_operator.Q.ptr_hst.data(), sizeof(int), (long long)subdomain_operator.Q.ptr_hst.size());
    case PRFDD_A_SUB_Q_COL: return put(subdomain_operator.Q.col_hst.data(), sizeof(int), (long long)subdomain_operator.Q.col_hst.size());
    case PRFDD_A_SUB_Q_VAL: return put(subdomain_operator.Q.val_hst.data(), sizeof(double), (long long)subdomain_operator.Q.val_hst.size());
    case PRFDD_A_SUB_ELEMENT_IDS:
    {
        std::vector<int> v;
        for (auto &e : subdomain_region) v.push_back(e.id);
        return put(v.data(), sizeof(int), (long long)v.size());
    }
    case PRFDD_A_SUB_ELEMENT_DEGREE:
    {
        std::vector<int> v;
        for (auto &e : subdomain_region) v.push_back(e.poly_degree);
        return put(v.data(), sizeof(int), (long long)v.size());
    }
    case PRFDD_A_SUB_DOF_NUM:
    {
        std::vector<long long> v;
        for (auto &e : subdomain_region) v.insert(v.end(), e.dof_num.begin(), e.dof_num.end());
        return put(v.data(), sizeof(long long), (long long)v.size());
    }
    case PRFDD_A_AMG_LEVEL_ROWS:
    {
        std::vector<int> v;
        for (auto &L : amg_fem.levels) v.push_back(L.n);
        return put(v.data(), sizeof(int), (long long)v.size());
    }
    case PRFDD_A_AMG_LEVEL_NNZ:
    {
        std::vector<int> v;
        for (auto &L : amg_fem.levels) v.push_back(L.A.nnz());
        return put(v.data(), sizeof(int), (long long)v.size());
    }
    case PRFDD_A_AMG_CHEBY_COEFS:
    {
        std::vector<double> v;
        for (auto &L : amg_fem.levels) v.insert(v.end(), L.coefs.begin(), L.coefs.end());
        return put(v.data(), sizeof(double), (long long)v.size());
    }
    case PRFDD_A_A_FEM_PTR: return put(A_fem_hst.ptr.data(), sizeof(int), (long long)A_fem_hst.ptr.size());
    case PRFDD_A_A_FEM_COL: return put(A_fem_hst.col.data(), sizeof(int), (long long)A_fem_hst.col.size());
    case PRFDD_A_A_FEM_VAL: return put(A_fem_hst.val.data(), sizeof(double), (long long)A_fem_hst.val.size());
    case PRFDD_A_NORM_WEIGHT: return put(norm_weight_hst.data(), sizeof(double), (long long)norm_weight_hst.size());
    case PRFDD_A_INNER_WEIGHT: return put(inner_weight_hst.data(), sizeof(double), (long long)inner_weight_hst.size());
    default: return -1;
    }
}

template <typename DType>
int Subdomain<DType>::apply(int what, const double *in_host, double *out_host)
{
    using namespace prfdd_host;
    const int nvl = num_values;
    switch (what)
    {
    case PRFDD_APPLY_SUB_STIFFNESS:
    {
        z_k.copyFrom(in_host, nvl * sizeof(double));
        stiffness_matrix(q_k, z_k);
        q_k.copyTo(out_host, nvl * sizeof(double));
        return 0;
    }
    case PRFDD_APPLY_LOW_ORDER:
    {
        z_k.copyFrom(in_host, nvl * sizeof(double));
        math.set_to_value(q_k, 0.0, nvl);
        low_order_preconditioner(q_k, z_k);
        q_k.copyTo(out_host, nvl * sizeof(double));
        return 0;
    }
    case PRFDD_APPLY_VCYCLE:
    {
        if (amg_fem.fp32)
        {
            work_dev[0].copyFrom(in_host, num_dofs * sizeof(double));
            dev::check_rc(prfdd_cast_f64_to_f32(amg_fem.levels[0].f.template as<float>(), dp(work_dev[0]), num_dofs, st()), "cast to FP32");
            amg_fem.vcycle(num_vcycles);
            dev::check_rc(prfdd_cast_f32_to_f64(dp(work_dev[0]), amg_fem.levels[0].u.template as<float>(), num_dofs, st()), "cast to FP64");
            work_dev[0].copyTo(out_host, num_dofs * sizeof(double));
            return 0;
        }
        amg_fem.levels[0].f.copyFrom(in_host, num_dofs * sizeof(double));
        amg_fem.vcycle(num_vcycles);
        amg_fem.levels[0].u.copyTo(out_host, num_dofs * sizeof(double));
        return 0;
    }
    case PRFDD_APPLY_TREE:
    {
        memory in = device.malloc<double>(own_points);
        in.copyFrom(in_host, own_points * sizeof(double));
        math.set_to_value(f, 0.0, nvl);
        tree_operator(f, in);
        f.copyTo(out_host, nvl * sizeof(double));
        return 0;
    }
    default: return -1;
    }
}

template <typename DType>
int Subdomain<DType>::time_spmv(int reps, double out[6])
{
    if (amg_fem.num_levels() < 2) return -1;
    cudaStream_t stream = st();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const bool f32 = amg_fem.fp32;
    auto step = [&](amg::Level &L) {
        if (f32) return prfdd_csrm_cheby_step_f32(L.u.as<float>(), L.t1.as<float>(), &L.dA.desc32, L.t0.as<float>(), L.r.as<float>(), L.ds.as<float>(), (float)L.coefs[0], 0, 0, stream);
        return prfdd_csrm_cheby_step(L.u.as<double>(), L.t1.as<double>(), &L.dA.desc, L.t0.as<double>(), L.r.as<double>(), L.ds.as<double>(), L.coefs[0], 0, 0, stream);
    };
    for (int w = 0; w < 4; w++) { step(amg_fem.levels[0]); step(amg_fem.levels[1]); }
    cudaEventRecord(e0, stream);
    for (int r = 0; r < reps; r++)
    {
        int rc = step(amg_fem.levels[r & 1]);
        if (rc) return rc;
    }
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double bytes = 0.0;
    for (int l = 0; l < 2; l++)
    {
        const amg::Level &L = amg_fem.levels[l];
        const double vb = f32 ? 4.0 : 8.0;
        bytes += 0.5 * ((4.0 + vb) * L.A.nnz() + 4.0 * (L.n + 1) + 4.0 * vb * L.n); // col + val per entry; ptr; t_in, ds, r read and t_out written once per row
        out[2 + 2 * l] = L.n;
        out[3 + 2 * l] = L.A.nnz();
    }
    out[0] = ms / reps;
    out[1] = bytes;
    return 0;
}

// one V-cycle with every launch bracketed by CUDA events (averaged over `reps` cycles): a text table, one line per launch,
// "level what rows nnz us algorithmic_MB GB/s"
template <typename DType>
int Subdomain<DType>::profile_vcycle(int reps, char *text, int cap)
{
    if (amg_fem.num_levels() < 1 || reps < 1) return -1;
    std::vector<std::vector<amg::Hierarchy::ProfileRec>> runs(reps);
    amg_fem.vcycle(num_vcycles); // warm
    for (int r = 0; r < reps; r++)
    {
        amg_fem.prof = &runs[r];
        amg_fem.vcycle(num_vcycles);
        amg_fem.prof = nullptr;
    }
    dev::check(cudaStreamSynchronize(st()), "profile_vcycle");
    std::string out;
    char line[256];
    double total_us = 0.0, total_bytes = 0.0;
    for (size_t k = 0; k < runs[0].size(); k++)
    {
        double us = 0.0;
        for (int r = 0; r < reps; r++)
        {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, runs[r][k].e0, runs[r][k].e1);
            us += 1e3 * ms / reps;
        }
        const auto &R = runs[0][k];
        snprintf(line, sizeof(line), "%d %-34s %9d %10d %8.2f %8.2f %8.1f\n", R.level, R.what, R.rows, R.nnz, us, R.bytes / 1e6, us > 0 ? R.bytes / us / 1e3 : 0.0);
        out += line;
        total_us += us;
        total_bytes += R.bytes;
    }
    snprintf(line, sizeof(line), "total: %zu launches, %.1f us (event-bracketed), %.1f MB algorithmic, %.1f GB/s\n", runs[0].size(), total_us, total_bytes / 1e6, total_bytes / total_us / 1e3);
    out += line;
    for (auto &run : runs)
        for (auto &R : run) { cudaEventDestroy(R.e0); cudaEventDestroy(R.e1); }
    if ((int)out.size() + 1 > cap) return -(int)out.size() - 1;
    memcpy(text, out.c_str(), out.size() + 1);
    return 0;
}

#include "subdomain_multi.hpp"
