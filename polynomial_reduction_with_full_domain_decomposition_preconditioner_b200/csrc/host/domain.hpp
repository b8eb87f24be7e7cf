// domain.hpp -- Domain<DType>: the outer Krylov solver and the fine matrix-free SEM operator.
//
// Same class surface as the reference (/root/reference/domain.hpp:35-145): initialize,
// initial_function, direct_stiffness_summation, stiffness_matrix, flexible_conjugate_gradient<P>,
// generalized_minimum_residual<P>; same public members and solver defaults (domain.hpp:112-118).
// What changed underneath (B200-first):
//   * stiffness_matrix is ONE fused launch (prfdd_stiffness_matrix), not two launches + 3 temporaries;
//   * Q / Q^T are applied as index maps (prfdd_gather / prfdd_scatter) on the CSR arrays the
//     reference builds (domain.tpp:286-294), with mask and 1/multiplicity folded in;
//   * the process-boundary sum (gslib_gs through host memory, domain.tpp:590-594) is a device-side
//     pack -> NCCL send/recv -> unpack-add in ascending-rank order;
//   * every reduction finishes on the device and feeds ncclAllReduce directly; alpha and beta are
//     formed on the device; the only host read per PCG iteration is the residual norm.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstring>
#include <map>
#include <string>
#include <typeinfo>
#include <unordered_map>
#include <vector>

#include "config.hpp"
#include "element.hpp"
#include "csr_matrix.hpp"
#include "math.hpp"
#include "special_functions.hpp"
#include "../../../include/prfdd_b200.h"

template <typename DType>
class Domain
{
    using memory = dev::memory;

  private:
    // Work arrays
    std::vector<std::vector<DType>> work_hst;
    std::vector<memory> work_dev;

    // Dirichlet boundary conditions
    memory dirichlet_mask;

    // Assembly
    CSR_Matrix<DType> Q;
    CSR_Matrix<DType> Qt;
    memory assembled_weight;

    // Gather scatter: process-boundary nodes are numbered first (domain.tpp:249-281)
    int num_bdary_nodes = 0;
    std::vector<long long> boundary_nodes;
    struct Halo
    {
        std::vector<int> peers;              // ascending rank
        std::vector<int> count;              // shared nodes with each peer
        std::vector<int> offset;             // into idx / buffers
        memory idx;                          // local node index of every (peer, shared node), sorted by global id
        memory send_buf, recv_buf, acc;      // acc: partial sum of lower-rank contributions
        int total = 0;
        int first_higher = 0;                // peers[first_higher..] have rank > proc_id
    } halo;

    // Solver
    memory r_k, r_kp1, q_k, z_k, p_k;
    std::vector<memory> V, Z, W;
    std::vector<std::vector<DType>> H;
    std::vector<DType> c_gmres, s_gmres, gamma;

    // device scalars + pinned host mirror
    memory scal;
    double *scal_hst = nullptr;
    prfdd_reduce_ws *ws = nullptr;

    Math<DType> math;

    double *dp(const memory &m) const { return m.as<double>(); }
    cudaStream_t st() const { return prfdd_host::device.stream; }

    void read_scalars(int first, int count)
    {
        dev::check(cudaMemcpyAsync(scal_hst + first, dp(scal) + first, count * sizeof(double), cudaMemcpyDeviceToHost, st()), "Domain::read_scalars");
        dev::check(cudaStreamSynchronize(st()), "Domain::read_scalars/sync");
    }

    void halo_exchange(const memory &nodes);
    void setup_halo();
    void residual_norm_dev(const memory &r); // the squared norm into scal[2], no host read
    template <typename PType> bool fcg_device_loop(memory &u, memory &f, PType &subdomain, bool use_relative);

  public:
    // Member variables
    std::string directory;
    int poly_degree = 0;
    const char *data_type = (typeid(DType) == typeid(double)) ? "double" : "float";

    int num_total_elements = 0;
    int num_total_points = 0;
    long long num_total_nodes = 0;

    int num_local_elements = 0;
    int num_local_points = 0;
    int num_local_nodes = 0;

    int num_elem_points = 0;

    // Elements
    std::vector<Element<DType>> elements;

    // Solver
    int num_blocks = 0;
    int num_iterations = 0;
    int num_vectors = 20;
    int max_iterations = 500;
    int preconditioner_type = 1;
    bool use_preconditioner = true;
    DType tolerance = (typeid(DType) == typeid(double)) ? 1.0e-07 : 1.0e-04;
    std::vector<double> history;
    // Outer FCG as ONE CUDA graph (the loop is a conditional WHILE node, its test runs on the device): no host round trip per
    // iteration.  The first solve on given buffers runs the host-driven loop (it creates every lazily built piece of state, the
    // preconditioner's own graphs among them); from the second solve on the graph is used.  Same kernels in the same order: the
    // residual history is identical bit for bit.
    bool device_outer_loop = false;
    struct OuterGraph
    {
        cudaGraphExec_t exec = nullptr;
        long long launches_head = 0, launches_body = 0;
        double bytes_head = 0.0, bytes_body = 0.0;
    };
    std::map<std::tuple<void *, void *, const void *, double, int, int>, OuterGraph> outer_graphs; // present with exec == nullptr: warmed, not built yet
    memory outer_hist, outer_state;
    double *outer_hist_hst = nullptr;
    int *outer_state_hst = nullptr;

    // Operator
    memory D_hat;
    std::vector<DType> D_hat_hst;
    memory geom_fact[NUM_GEOM_FACTS];

    Domain() {}
    Domain(const char *directory_, int poly_degree_) { initialize(directory_, poly_degree_); }
    ~Domain()
    {
        if (ws) prfdd_reduce_ws_destroy(ws);
        if (scal_hst) cudaFreeHost(scal_hst);
    }
    Domain(const Domain &) = delete;
    Domain &operator=(const Domain &) = delete;

    void initialize(const char *directory_, int poly_degree_, bool solver_buffers = true);

    // Member functions
    void initial_function(memory &u, int function_id = 0);
    void direct_stiffness_summation(const memory &QQtu, const memory &u, bool apply_dirichlet_mask = true, bool apply_assembled_weight = false);
    void stiffness_matrix(const memory &Au, const memory &u, bool apply_dssum = false);

    template <typename PType>
    void flexible_conjugate_gradient(memory &u, memory &f, PType &subdomain, bool use_relative = true);

    template <typename PType>
    void generalized_minimum_residual(memory &u, memory &f, PType &subdomain, bool use_relative = true);

    // accessors for the C ABI / tests
    const CSR_Matrix<DType> &Q_matrix() const { return Q; }
    const CSR_Matrix<DType> &Qt_matrix() const { return Qt; }
    const std::vector<long long> &boundary_node_ids() const { return boundary_nodes; }
    int num_boundary_nodes() const { return num_bdary_nodes; }
    const memory &assembled_weight_dev() const { return assembled_weight; }
    const memory &dirichlet_mask_dev() const { return dirichlet_mask; }
    prfdd_reduce_ws *reduce_ws() const { return ws; }

  private:
    template <typename PType>
    void precondition(memory &z, memory &r, PType &subdomain);
    DType residual_norm_sync(const memory &r);
};

// ---------------------------------------------------------------------------------------------
// implementation
// ---------------------------------------------------------------------------------------------
namespace prfdd_detail
{
template <typename T>
inline bool read_block(const std::string &path, T *dst, size_t count)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    size_t got = fread(dst, sizeof(T), count, f);
    fclose(f);
    return got == count;
}
} // namespace prfdd_detail

template <typename DType>
void Domain<DType>::initialize(const char *directory_, int poly_degree_, bool solver_buffers)
{
    using namespace prfdd_host;
    directory = directory_;
    poly_degree = poly_degree_;
    setup_submark(nullptr);

    char file_name[4096];
    auto path = [&](const char *name) {
        snprintf(file_name, sizeof(file_name), "%s/lx1_%d/%s_%d.%d.dat", directory.c_str(), poly_degree + 1, name, proc_id, poly_degree);
        return std::string(file_name);
    };

    // Size data (domain.tpp:43-54); `dim` is a side effect of reading the size file (tpp:47)
    {
        int n_x, n_y, n_z;
        FILE *file_ptr = fopen(path("size").c_str(), "r");
        if (!file_ptr) throw std::runtime_error("ERROR: cannot open " + path("size"));
        int got = fscanf(file_ptr, "%d %d %d %d %d", &dim, &n_x, &n_y, &n_z, &num_local_elements);
        fclose(file_ptr);
        if (got != 5) throw std::runtime_error("ERROR: There was a problem reading Nek5000 data (size file)");
    }
    num_total_elements = (int)comm_world.allreduce_sum_host(num_local_elements);
    num_elem_points = 1;
    for (int d = 0; d < dim; d++) num_elem_points *= (poly_degree + 1);
    num_local_points = num_local_elements * num_elem_points;
    num_total_points = num_total_elements * num_elem_points;
    const int P = num_local_points;

    // Work arrays (tpp:56-67)
    work_hst.resize(dim);
    for (auto &w : work_hst) w.resize(P);
    work_dev.resize(dim);
    for (auto &w : work_dev) w = device.malloc<DType>(P);

    // Elements (tpp:69-77)
    elements.clear();
    elements.reserve(num_local_elements);
    for (int e = 0; e < num_local_elements; e++)
    {
        elements.push_back(Element<DType>(e, dim, poly_degree));
        if (e > 0) elements[e].offset = elements[e - 1].offset + elements[e].num_points;
    }

    bool ok = true;
    std::vector<DType> buf(P);
    auto read_per_element = [&](const char *name, auto member) {
        ok = ok && prfdd_detail::read_block(path(name), buf.data(), (size_t)P);
        if (!ok) return;
        for (auto &elem : elements) memcpy(member(elem).data(), buf.data() + elem.offset, elem.num_points * sizeof(DType));
    };
    // Geometry (tpp:79-138)
    if (dim >= 1) read_per_element("x", [](Element<DType> &e) -> std::vector<DType> & { return e.x; });
    if (dim >= 2) read_per_element("y", [](Element<DType> &e) -> std::vector<DType> & { return e.y; });
    if (dim >= 3) read_per_element("z", [](Element<DType> &e) -> std::vector<DType> & { return e.z; });

    // Connectivity (tpp:140-160)
    {
        std::vector<long long> gbuf(P);
        ok = ok && prfdd_detail::read_block(path("glo_num"), gbuf.data(), (size_t)P);
        if (ok)
            for (auto &elem : elements) memcpy(elem.glo_num.data(), gbuf.data() + elem.offset, elem.num_points * sizeof(long long));
        for (auto &elem : elements)
            for (int v = 0; v < elem.num_points; v++) elem.loc_num[v] = elem.offset + v;
    }

    // Node degree (tpp:162-172)
    std::vector<int> node_degree(P);
    ok = ok && prfdd_detail::read_block(path("node_degree"), node_degree.data(), (size_t)P);

    // Dirichlet boundary conditions (tpp:174-194)
    read_per_element("p_mask", [](Element<DType> &e) -> std::vector<DType> & { return e.dirichlet_mask; });
    dirichlet_mask = device.malloc<DType>(P);
    if (ok) dirichlet_mask.copyFrom(buf.data(), P * sizeof(DType));

    // Geometric factors (tpp:196-224); all six files are read, also in 2D
    for (int g = 0; g < NUM_GEOM_FACTS; g++)
    {
        char name[16];
        snprintf(name, sizeof(name), "g_%d", g + 1);
        ok = ok && prfdd_detail::read_block(path(name), buf.data(), (size_t)P);
        if (!ok) break;
        for (auto &elem : elements) memcpy(elem.geom_fact[g].data(), buf.data() + elem.offset, elem.num_points * sizeof(DType));
        geom_fact[g] = device.malloc<DType>(P);
        geom_fact[g].copyFrom(buf.data(), P * sizeof(DType));
    }

    if (!ok)
    {
        pstdout("ERROR: There was a problem reading Nek5000 data\n");
        throw std::runtime_error("ERROR: There was a problem reading Nek5000 data in " + directory);
    }

    setup_submark("read mesh files, upload geometry");
    // Communication (tpp:233-284)
    if (proc_id == 0 && verbose) printf("Setting up domain stitching handle...\n");

    std::unordered_map<long long, int> local_node_degree;
    local_node_degree.reserve((size_t)P);
    for (auto &elem : elements)
        for (int v = 0; v < elem.num_points; v++) local_node_degree[elem.glo_num[v]]++;

    std::unordered_map<long long, int> local_node_idx;
    local_node_idx.reserve(local_node_degree.size());
    boundary_nodes.clear();
    int count = 0;
    for (auto &elem : elements)
        for (int v = 0; v < elem.num_points; v++)
            if (local_node_degree[elem.glo_num[v]] != node_degree[elem.offset + v])
                if (local_node_idx.find(elem.glo_num[v]) == local_node_idx.end())
                {
                    boundary_nodes.push_back(elem.glo_num[v]);
                    local_node_idx[elem.glo_num[v]] = count;
                    count++;
                }
    num_bdary_nodes = count;
    for (auto &elem : elements)
        for (int v = 0; v < elem.num_points; v++)
            if (local_node_idx.find(elem.glo_num[v]) == local_node_idx.end())
            {
                local_node_idx[elem.glo_num[v]] = count;
                count++;
            }

    setup_submark("node numbering (boundary first, first touch)");
    setup_halo(); // stands where gslib_gs_setup stands (tpp:283-284)
    setup_submark("halo lists");

    // Q, Qt (tpp:286-294)
    num_local_nodes = (int)local_node_degree.size();
    Q.initialize(P, num_local_nodes);
    Q.reserve(P);
    for (auto &elem : elements)
        for (int v = 0; v < elem.num_points; v++) Q.add_entry(elem.loc_num[v], local_node_idx[elem.glo_num[v]], 1.0);
    Q.assemble();
    Q.transpose(Qt);
    setup_submark("Q, Qt");

    // total number of unique nodes = sum over ranks and nodes of 1/multiplicity... counted exactly with integers below
    // assembled_weight = 1 / (Qt 1 (+) gs_add)  (tpp:296-302)
    assembled_weight = device.malloc<DType>(std::max(num_local_nodes, 1));
    math.set_to_value(work_dev[0], 1.0, P);
    dev::check_rc(prfdd_gather(dp(assembled_weight), Qt.ptr.template as<int>(), Qt.col.template as<int>(), dp(work_dev[0]), nullptr, num_local_nodes, st()), "Domain::initialize/gather");
    halo_exchange(assembled_weight);
    math.invert_vector_elements(assembled_weight, num_local_nodes);

    // number of global unique nodes: sum_v 1/mult_v over all ranks, done in integers via lcm-free counting:
    // every node is counted by the lowest rank that holds it
    {
        std::vector<DType> w(num_local_nodes);
        assembled_weight.copyTo(w.data(), num_local_nodes * sizeof(DType));
        double s = 0.0;
        // multiplicity m => this rank holds local_mult of the m copies
        std::vector<int> local_mult(num_local_nodes, 0);
        for (int i = 0; i < P; i++) local_mult[Q.col_hst[i]]++;
        for (int v = 0; v < num_local_nodes; v++) s += w[v] * local_mult[v];
        scal = device.malloc<double>(16);
        dev::check(cudaMallocHost((void **)&scal_hst, 16 * sizeof(double)), "Domain/pinned scalars");
        scal_hst[15] = s;
        scal.copyFrom(scal_hst, 16 * sizeof(double));
        comm_world.allreduce_sum(dp(scal) + 15, 1);
        read_scalars(15, 1);
        num_total_nodes = (long long)std::llround(scal_hst[15]);
    }

    // Operator (tpp:304-316)
    {
        int num_gll_points = poly_degree + 1;
        std::vector<double> r_gll(num_gll_points), w_gll(num_gll_points);
        std::vector<double> D_gll(num_gll_points * num_gll_points), Dt_gll(num_gll_points * num_gll_points);
        zwgll_(r_gll.data(), w_gll.data(), &num_gll_points);
        dgll_(Dt_gll.data(), D_gll.data(), r_gll.data(), &num_gll_points, &num_gll_points);
        D_hat_hst.assign(D_gll.begin(), D_gll.end());
        D_hat = device.malloc<DType>(num_gll_points * num_gll_points);
        D_hat.copyFrom(D_hat_hst.data(), num_gll_points * num_gll_points * sizeof(DType));
    }

    num_blocks = (P + BLOCK_SIZE - 1) / BLOCK_SIZE;
    dev::check_rc(prfdd_reduce_ws_create(&ws), "prfdd_reduce_ws_create");

    // Solver (tpp:318-330).  The 41 outer-GMRES vectors are allocated lazily, on first use.
    if (solver_buffers)
    {
        r_k = device.malloc<DType>(P);
        r_kp1 = device.malloc<DType>(P);
        q_k = device.malloc<DType>(P);
        z_k = device.malloc<DType>(P);
        p_k = device.malloc<DType>(P);
    }
    setup_submark("weights, node count, buffers");
    H.assign(num_vectors, std::vector<DType>(num_vectors));
    c_gmres.assign(num_vectors, 0);
    s_gmres.assign(num_vectors, 0);
    gamma.assign(num_vectors + 1, 0);
}

template <typename DType>
void Domain<DType>::setup_halo()
{
    using namespace prfdd_host;
    halo = Halo();
    if (num_procs == 1) return;
    // every rank learns every rank's boundary ids
    long long my_n = num_bdary_nodes;
    long long max_n = comm_world.allreduce_max_host(my_n);
    std::vector<long long> mine(max_n + 1, 0), all((size_t)(max_n + 1) * num_procs);
    mine[0] = my_n;
    for (long long i = 0; i < my_n; i++) mine[1 + i] = boundary_nodes[i];
    comm_world.allgather_host(mine.data(), all.data(), (size_t)(max_n + 1) * sizeof(long long));

    // flatten to (ids, offsets) and let the shared host routine build the per-peer lists
    std::vector<long long> ids, offsets(num_procs + 1, 0);
    for (int p = 0; p < num_procs; p++)
    {
        const long long *rec = all.data() + (size_t)p * (max_n + 1);
        offsets[p + 1] = offsets[p] + rec[0];
        ids.insert(ids.end(), rec + 1, rec + 1 + rec[0]);
    }
    std::vector<int> peers(num_procs), pcount(num_procs), poffset(num_procs), idx_all((size_t)std::max<long long>(my_n, 1) * num_procs);
    long long total = 0;
    int np = prfdd_halo_build_lists(proc_id, num_procs, ids.data(), offsets.data(), peers.data(), pcount.data(), poffset.data(), idx_all.data(), (long long)idx_all.size(), &total);
    if (np < 0) throw std::runtime_error("Domain::setup_halo: prfdd_halo_build_lists failed");
    halo.peers.assign(peers.begin(), peers.begin() + np);
    halo.count.assign(pcount.begin(), pcount.begin() + np);
    halo.offset.assign(poffset.begin(), poffset.begin() + np);
    idx_all.resize(total);
    halo.first_higher = 0;
    while (halo.first_higher < (int)halo.peers.size() && halo.peers[halo.first_higher] < proc_id) halo.first_higher++;
    halo.total = (int)idx_all.size();
    halo.idx = device.malloc<int>(std::max(halo.total, 1));
    halo.idx.copyFrom(idx_all.data(), halo.total * sizeof(int));
    halo.send_buf = device.malloc<double>(std::max(halo.total, 1));
    halo.recv_buf = device.malloc<double>(std::max(halo.total, 1));
    halo.acc = device.malloc<double>(std::max(num_bdary_nodes, 1));
}

// gs_add over the process-boundary nodes (the first num_bdary_nodes entries of a node vector).
// The sum at every shared node is formed in ascending-rank order on every holder, so all holders end
// up with bit-identical values.
template <typename DType>
void Domain<DType>::halo_exchange(const memory &nodes)
{
    using namespace prfdd_host;
    if (num_procs == 1 || halo.peers.empty()) return;
    double *nd = nodes.as<double>();
    const int *idx = halo.idx.template as<int>();
    dev::check_rc(prfdd_halo_pack(dp(halo.send_buf), nd, idx, halo.total, st()), "halo_pack");
    std::vector<const double *> sp;
    std::vector<double *> rp;
    std::vector<size_t> cnt;
    for (size_t k = 0; k < halo.peers.size(); k++)
    {
        sp.push_back(dp(halo.send_buf) + halo.offset[k]);
        rp.push_back(dp(halo.recv_buf) + halo.offset[k]);
        cnt.push_back((size_t)halo.count[k]);
    }
    comm_world.sendrecv(halo.peers, sp, cnt, rp, cnt);
    if (halo.first_higher > 0)
    {
        // lower ranks first: acc = sum_{q<me} v_q ; nodes = acc + nodes
        dev::check_rc(prfdd_set_to_value(dp(halo.acc), 0.0, num_bdary_nodes, 0, st()), "halo acc");
        for (int k = 0; k < halo.first_higher; k++)
            dev::check_rc(prfdd_halo_unpack_add(dp(halo.acc), dp(halo.recv_buf) + halo.offset[k], idx + halo.offset[k], halo.count[k], st()), "halo_unpack_add");
        dev::check_rc(prfdd_vector_vector_addition(nd, 1.0, dp(halo.acc), 1.0, nd, num_bdary_nodes, st()), "halo combine");
    }
    for (size_t k = halo.first_higher; k < halo.peers.size(); k++)
        dev::check_rc(prfdd_halo_unpack_add(nd, dp(halo.recv_buf) + halo.offset[k], idx + halo.offset[k], halo.count[k], st()), "halo_unpack_add");
}

template <typename DType>
void Domain<DType>::initial_function(memory &u, int function_id)
{
    using namespace prfdd_host;
    // domain.tpp:527-580
    std::vector<DType> &w = work_hst[0];
    if (function_id == 4)
    {
        // glibc rand() with the default seed, drawn in element-major point order, per rank (tpp:549-550, 572-573)
        prfdd_glibc_rand_fill(w.data(), num_local_points, 1u);
    }
    else
    {
        for (auto &elem : elements)
            for (int v = 0; v < elem.num_points; v++)
            {
                const double sx = sin(M_PI * elem.x[v]), sy = sin(M_PI * elem.y[v]);
                const double sz = (dim == 2) ? 1.0 : sin(M_PI * elem.z[v]);
                double val = 0.0;
                if (function_id == 0) val = sx * sy * sz;
                else if (function_id == 1) val = sx * sy * sz + sin(2.0 * M_PI * elem.x[v]) * sy * sz;
                else if (function_id == 2) val = exp(elem.x[v]) * sx * sy * sz;
                else throw std::runtime_error("initial_function: function_id not supported");
                w[elem.loc_num[v]] = val;
            }
    }
    u.copyFrom(w.data(), num_local_points * sizeof(DType));
    direct_stiffness_summation(u, u, true, true);
}

template <typename DType>
void Domain<DType>::direct_stiffness_summation(const memory &QQtu, const memory &u, bool apply_dirichlet_mask, bool apply_assembled_weight)
{
    // domain.tpp:582-600
    dev::check_rc(prfdd_gather(dp(work_dev[0]), Qt.ptr.template as<int>(), Qt.col.template as<int>(), dp(u), apply_assembled_weight ? dp(assembled_weight) : nullptr, num_local_nodes, st()), "dssum/gather");
    prfdd_algorithmic_bytes_add(12.0 * num_local_points + 8.0 * num_local_nodes); // gathered points (index + value) and the node vector the scatter reads
    halo_exchange(work_dev[0]);
    dev::check_rc(prfdd_scatter(dp(QQtu), Q.col.template as<int>(), dp(work_dev[0]), apply_dirichlet_mask ? dp(dirichlet_mask) : nullptr, num_local_points, st()), "dssum/scatter");
}

template <typename DType>
void Domain<DType>::stiffness_matrix(const memory &Au, const memory &u, bool apply_dssum)
{
    // domain.tpp:602-609
    const double *g[6];
    for (int c = 0; c < 6; c++) g[c] = dp(geom_fact[c]);
    dev::check_rc(prfdd_stiffness_matrix_hd(dp(Au), dp(u), dp(D_hat), D_hat_hst.data(), g, num_local_elements, poly_degree + 1, prfdd_host::dim, st()), "Domain::stiffness_matrix");
    if (apply_dssum) direct_stiffness_summation(Au, Au, true, false);
}

template <typename DType>
DType Domain<DType>::residual_norm_sync(const memory &r)
{
    // domain.tpp:916-931: sqrt( allreduce( sum r * QQt r * mask ) )
    residual_norm_dev(r);
    read_scalars(2, 1);
    return std::sqrt(scal_hst[2]);
}

template <typename DType>
void Domain<DType>::residual_norm_dev(const memory &r)
{
    direct_stiffness_summation(work_dev[1], r);
    dev::check_rc(prfdd_residual_norm(ws, dp(scal) + 2, dp(r), dp(work_dev[1]), dp(dirichlet_mask), num_local_points, st()), "residual_norm");
    prfdd_host::comm_world.allreduce_sum(dp(scal) + 2, 1);
}

// Outer FCG as one graph.  Returns false when this call has to take the host-driven loop (first solve on these buffers, timers on).
// The loop of domain.tpp:621-725 is rotated so that its test sits at the end of the WHILE body:
//   head:  initialise, |r_0|, z = M r, p = z, FIRST HALF(0), test
//   body:  SECOND HALF (z = M r+, beta, p, r), FIRST HALF (q = A p, gamma, theta, u, r+, |r+|), test
// which executes exactly the launches of the host loop up to its `break`.
template <typename DType>
template <typename PType>
bool Domain<DType>::fcg_device_loop(memory &u, memory &f, PType &subdomain, bool use_relative)
{
    using namespace prfdd_host;
    if (!device_outer_loop || timer.enabled || max_iterations < 1) return false;
    const auto key = std::make_tuple(u.ptr(), f.ptr(), (const void *)&subdomain, (double)tolerance, use_relative ? 1 : 0, max_iterations);
    auto it = outer_graphs.find(key);
    if (it == outer_graphs.end())
    {
        outer_graphs.emplace(key, OuterGraph()); // warmed by the host-driven solve the caller runs now
        return false;
    }
    const int P = num_local_points;
    double *sc = dp(scal);
    memory &u_k = u;
    if (!outer_hist.is_initialized())
    {
        outer_hist = device.malloc<double>(max_iterations + 2);
        outer_state = device.malloc<int>(2);
        dev::check(cudaMallocHost((void **)&outer_hist_hst, sizeof(double) * (max_iterations + 2)), "pinned history");
        dev::check(cudaMallocHost((void **)&outer_state_hst, sizeof(int) * 2), "pinned loop state");
    }
    double *hist = outer_hist.template as<double>();
    int *state = outer_state.template as<int>();
    OuterGraph &og = it->second;
    if (!og.exec)
    {
        cudaGraph_t g = nullptr;
        dev::check(cudaGraphCreate(&g, 0), "cudaGraphCreate");
        cudaGraphConditionalHandle handle;
        dev::check(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault), "cudaGraphConditionalHandleCreate");
        auto first_half = [&](const memory &r_out) {
            stiffness_matrix(q_k, p_k);
            dev::check_rc(prfdd_projection_inner_products(ws, sc + 0, dp(z_k), dp(r_k), dp(p_k), dp(q_k), P, st()), "projection_inner_products");
            comm_world.allreduce_sum(sc + 0, 2);
            dev::check_rc(prfdd_solution_and_residual_update_dev(dp(u_k), dp(r_out), dp(r_k), dp(p_k), dp(q_k), sc + 0, sc + 1, P, st()), "solution_and_residual_update");
            residual_norm_dev(r_out);
            dev::check_rc(prfdd_fcg_outer_check(sc + 2, hist, state, (double)tolerance, use_relative ? 1 : 0, max_iterations, (unsigned long long)handle, 1, st()), "fcg_outer_check");
        };
        // head
        long long l0 = prfdd_launch_count();
        double b0 = prfdd_algorithmic_bytes();
        dev::check(cudaStreamBeginCaptureToGraph(st(), g, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCaptureToGraph(head)");
        dev::check_rc(prfdd_fcg_outer_reset(state, st()), "fcg_outer_reset");
        dev::check_rc(prfdd_initialize_arrays(dp(u_k), dp(r_k), dp(f), P, st()), "initialize_arrays");
        residual_norm_dev(r_k);
        dev::check_rc(prfdd_fcg_outer_check(sc + 2, hist, state, (double)tolerance, use_relative ? 1 : 0, max_iterations, 0ull, 0, st()), "fcg_outer_check(r0)");
        precondition(z_k, r_k, subdomain);
        p_k.copyFrom(z_k, P * sizeof(DType));
        first_half(r_kp1);
        cudaGraph_t same = nullptr;
        dev::check(cudaStreamEndCapture(st(), &same), "cudaStreamEndCapture(head)");
        og.launches_head = prfdd_launch_count() - l0;
        og.bytes_head = prfdd_algorithmic_bytes() - b0;
        // the WHILE node after every leaf of the head
        size_t n_nodes = 0;
        dev::check(cudaGraphGetNodes(g, nullptr, &n_nodes), "cudaGraphGetNodes");
        std::vector<cudaGraphNode_t> nodes(n_nodes), leaves;
        dev::check(cudaGraphGetNodes(g, nodes.data(), &n_nodes), "cudaGraphGetNodes");
        for (auto nd : nodes)
        {
            size_t n_dep = 0;
            dev::check(cudaGraphNodeGetDependentNodes(nd, nullptr, &n_dep), "cudaGraphNodeGetDependentNodes");
            if (n_dep == 0) leaves.push_back(nd);
        }
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.conditional.handle = handle;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        cudaGraphNode_t cnode;
        dev::check(cudaGraphAddNode(&cnode, g, leaves.data(), leaves.size(), &cp), "cudaGraphAddNode(conditional)");
        cudaGraph_t body = cp.conditional.phGraph_out[0];
        l0 = prfdd_launch_count();
        b0 = prfdd_algorithmic_bytes();
        dev::check(cudaStreamBeginCaptureToGraph(st(), body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCaptureToGraph(body)");
        precondition(z_k, r_kp1, subdomain);
        dev::check_rc(prfdd_inner_product_flexible(ws, sc + 3, dp(r_k), dp(r_kp1), dp(z_k), P, st()), "inner_product_flexible");
        comm_world.allreduce_sum(sc + 3, 1);
        dev::check_rc(prfdd_residual_and_search_update_dev(dp(p_k), dp(r_k), dp(z_k), dp(r_kp1), sc + 3, sc + 0, P, st()), "residual_and_search_update");
        dev::check_rc(prfdd_fcg_outer_count(state, st()), "fcg_outer_count");
        first_half(r_kp1);
        dev::check(cudaStreamEndCapture(st(), &same), "cudaStreamEndCapture(body)");
        og.launches_body = prfdd_launch_count() - l0;
        og.bytes_body = prfdd_algorithmic_bytes() - b0;
        prfdd_launch_count_add(-(og.launches_head + og.launches_body)); // capturing launched nothing
        prfdd_algorithmic_bytes_add(-(og.bytes_head + og.bytes_body));
        dev::check(cudaGraphInstantiate(&og.exec, g, 0), "cudaGraphInstantiate(outer FCG)");
        cudaGraphDestroy(g);
    }
    dev::check(cudaGraphLaunch(og.exec, st()), "cudaGraphLaunch(outer FCG)");
    dev::check(cudaMemcpyAsync(outer_state_hst, state, 2 * sizeof(int), cudaMemcpyDeviceToHost, st()), "loop state");
    dev::check(cudaMemcpyAsync(outer_hist_hst, hist, (max_iterations + 2) * sizeof(double), cudaMemcpyDeviceToHost, st()), "history");
    dev::check(cudaStreamSynchronize(st()), "outer FCG graph");
    const int norms = outer_state_hst[0], updates = outer_state_hst[1];
    history.assign(outer_hist_hst, outer_hist_hst + norms);
    num_iterations = updates;
    const double r0 = history[0], last = history.back();
    const bool converged = use_relative ? (last / r0 < tolerance) : (last < tolerance);
    if (!converged && !std::isnan(last) && norms - 1 >= max_iterations) num_iterations = max_iterations; // the host loop's last search update changes neither u nor the history
    prfdd_launch_count_add(og.launches_head + (long long)updates * og.launches_body);
    prfdd_algorithmic_bytes_add(og.bytes_head + updates * og.bytes_body);
    for (int k = 0; k < norms; k++)
        rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | \n", k, history[k], history[k] / r0);
    return true;
}

template <typename DType>
template <typename PType>
void Domain<DType>::precondition(memory &z, memory &r, PType &subdomain)
{
    using namespace prfdd_host;
    // domain.tpp:637-651, 697-711
    if (use_preconditioner)
    {
        if (preconditioner_type == 0)
            subdomain.flexible_conjugate_gradient(z, r);
        else
            subdomain.generalized_minimum_residual(z, r);
        timer.start("subdomain.stitching");
        direct_stiffness_summation(z, z, true, true);
        timer.stop("subdomain.stitching");
    }
    else
    {
        direct_stiffness_summation(z, r);
    }
}

template <typename DType>
template <typename PType>
void Domain<DType>::flexible_conjugate_gradient(memory &u, memory &f, PType &subdomain, bool use_relative)
{
    using namespace prfdd_host;
    if (fcg_device_loop(u, f, subdomain, use_relative)) return;
    const int P = num_local_points;
    // device scalars: [0] gamma  [1] theta  [2] |r|^2  [3] theta_flex
    double *sc = dp(scal);

    // Initialize arrays (domain.tpp:616-619)
    timer.start("domain.vector_operations");
    memory &u_k = u;
    dev::check_rc(prfdd_initialize_arrays(dp(u_k), dp(r_k), dp(f), P, st()), "initialize_arrays");
    timer.stop("domain.vector_operations");

    // Compute initial residual
    timer.start("domain.residual_norm");
    DType r_0_norm = residual_norm_sync(r_k);
    DType r_norm = r_0_norm;
    timer.stop("domain.residual_norm");
    history.assign(1, r_0_norm);
    rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | \n", 0, r_0_norm, 1.0);

    precondition(z_k, r_k, subdomain);

    timer.start("domain.vector_operations");
    p_k.copyFrom(z_k, P * sizeof(DType));
    timer.stop("domain.vector_operations");

    num_iterations = 0;

    for (int iter = 0; iter < max_iterations; iter++)
    {
        // Projection
        timer.start("domain.operator_application");
        stiffness_matrix(q_k, p_k);
        timer.stop("domain.operator_application");

        // Inner products: gamma = z.r, theta = p.q   (stay on the device)
        timer.start("domain.inner_products");
        dev::check_rc(prfdd_projection_inner_products(ws, sc + 0, dp(z_k), dp(r_k), dp(p_k), dp(q_k), P, st()), "projection_inner_products");
        comm_world.allreduce_sum(sc + 0, 2);
        timer.stop("domain.inner_products");

        // Update solution and residual with alpha = gamma / theta formed on the device
        timer.start("domain.vector_operations");
        dev::check_rc(prfdd_solution_and_residual_update_dev(dp(u_k), dp(r_kp1), dp(r_k), dp(p_k), dp(q_k), sc + 0, sc + 1, P, st()), "solution_and_residual_update");
        timer.stop("domain.vector_operations");

        // Residual norm: the one host read of the iteration
        timer.start("domain.residual_norm");
        r_norm = residual_norm_sync(r_kp1);
        timer.stop("domain.residual_norm");
        history.push_back(r_norm);

        rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | \n", iter + 1, r_norm, r_norm / r_0_norm);

        if (use_relative)
        {
            if (r_norm / r_0_norm < tolerance) break;
        }
        else
        {
            if (r_norm < tolerance) break;
        }
        if (std::isnan(r_norm)) break;

        // Update search direction
        precondition(z_k, r_kp1, subdomain);

        timer.start("domain.inner_products");
        dev::check_rc(prfdd_inner_product_flexible(ws, sc + 3, dp(r_k), dp(r_kp1), dp(z_k), P, st()), "inner_product_flexible");
        comm_world.allreduce_sum(sc + 3, 1);
        timer.stop("domain.inner_products");

        // beta = theta_flex / gamma on the device
        timer.start("domain.vector_operations");
        dev::check_rc(prfdd_residual_and_search_update_dev(dp(p_k), dp(r_k), dp(z_k), dp(r_kp1), sc + 3, sc + 0, P, st()), "residual_and_search_update");
        timer.stop("domain.vector_operations");

        num_iterations++;
    }
}

template <typename DType>
template <typename PType>
void Domain<DType>::generalized_minimum_residual(memory &u, memory &f, PType &subdomain, bool use_relative)
{
    using namespace prfdd_host;
    const int P = num_local_points;
    double *sc = dp(scal);

    // the reference allocates V[21], Z[20] in initialize() (domain.tpp:325-326); here on first use.
    // W[i] = mask .* QQt V[i] is cached so that the j+1 Gram-Schmidt dots of a column are one pass.
    if (V.empty())
    {
        V.resize(num_vectors + 1);
        for (auto &v : V) v = device.malloc<DType>(P);
        Z.resize(num_vectors);
        for (auto &z : Z) z = device.malloc<DType>(P);
        W.resize(num_vectors + 1);
        for (auto &w : W) w = device.malloc<DType>(P);
    }
    memory hcol = device.malloc<double>(num_vectors + 1);
    std::vector<double> hcol_hst(num_vectors + 1);

    timer.start("domain.vector_operations");
    memory &u_k = u;
    dev::check_rc(prfdd_initialize_arrays(dp(u_k), dp(r_k), dp(f), P, st()), "initialize_arrays");
    timer.stop("domain.vector_operations");

    timer.start("domain.residual_norm");
    DType r_0_norm = residual_norm_sync(r_k);
    DType r_norm = r_0_norm;
    timer.stop("domain.residual_norm");
    history.assign(1, r_0_norm);
    rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | \n", 0, r_0_norm, 1.0);

    bool converged = false;
    int iter = 0;
    int j;
    DType alpha_j, beta_j, gamma_j, gamma_k;

    while (iter < max_iterations)
    {
        if (iter > 0)
        {
            timer.start("domain.operator_application");
            stiffness_matrix(r_k, u_k);
            timer.stop("domain.operator_application");
            math.vector_vector_addition(r_k, 1.0, f, -1.0, r_k, P);
            timer.start("domain.residual_norm");
            r_norm = residual_norm_sync(r_k);
            timer.stop("domain.residual_norm");
            gamma[0] = r_norm;
        }
        else
        {
            gamma[0] = r_0_norm;
        }

        math.vector_scaling(V[0], 1.0 / gamma[0], r_k, P);
        direct_stiffness_summation(W[0], V[0]); // mask applied in the scatter

        for (j = 0; j < num_vectors; j++)
        {
            precondition(Z[j], V[j], subdomain);

            timer.start("domain.operator_application");
            stiffness_matrix(q_k, Z[j]);
            timer.stop("domain.operator_application");

            // Gram-Schmidt, first pass (domain.tpp:810-822): H[i][j] = sum q * (QQt V_i) * mask
            timer.start("domain.inner_products");
            {
                std::vector<const double *> wp(j + 1);
                for (int i = 0; i < j + 1; i++) wp[i] = dp(W[i]);
                dev::check_rc(prfdd_multi_inner_product(ws, dp(hcol), dp(q_k), wp.data(), nullptr, j + 1, P, st()), "multi_inner_product");
                comm_world.allreduce_sum(dp(hcol), j + 1);
                hcol.copyTo(hcol_hst.data(), (j + 1) * sizeof(double));
                for (int i = 0; i < j + 1; i++) H[i][j] = hcol_hst[i];
            }
            timer.stop("domain.inner_products");

            timer.start("domain.vector_operations");
            {
                std::vector<const double *> vp(j + 1);
                for (int i = 0; i < j + 1; i++) vp[i] = dp(V[i]);
                dev::check_rc(prfdd_multi_axpy_dev(dp(q_k), vp.data(), dp(hcol), 1, -1.0, j + 1, P, st()), "multi_axpy");
            }
            timer.stop("domain.vector_operations");

            // Givens rotations on the new column
            for (int i = 0; i < j; i++)
            {
                DType h_ij = H[i][j];
                H[i][j] = c_gmres[i] * h_ij + s_gmres[i] * H[i + 1][j];
                H[i + 1][j] = -s_gmres[i] * h_ij + c_gmres[i] * H[i + 1][j];
            }

            timer.start("domain.residual_norm");
            // residual_norm(q) also yields QQt q = alpha_j * QQt V[j+1]: keep it for W[j+1]
            direct_stiffness_summation(W[j + 1], q_k);
            dev::check_rc(prfdd_residual_norm(ws, sc + 2, dp(q_k), dp(W[j + 1]), dp(dirichlet_mask), P, st()), "residual_norm");
            comm_world.allreduce_sum(sc + 2, 1);
            read_scalars(2, 1);
            alpha_j = std::sqrt(scal_hst[2]);
            timer.stop("domain.residual_norm");

            if (std::abs(alpha_j) == 0.0) { converged = true; break; }

            beta_j = std::sqrt(H[j][j] * H[j][j] + alpha_j * alpha_j);
            gamma_j = 1.0 / beta_j;
            c_gmres[j] = H[j][j] * gamma_j;
            s_gmres[j] = alpha_j * gamma_j;
            H[j][j] = beta_j;
            gamma[j + 1] = -s_gmres[j] * gamma[j];
            gamma[j] = c_gmres[j] * gamma[j];

            r_norm = std::abs(gamma[j + 1]);
            history.push_back(r_norm);
            rstdout("Iter %2d: | residual_norm = %24.16g | relative_residual_norm = %24.16g | \n", iter + 1, r_norm, r_norm / r_0_norm);

            if (use_relative ? (r_norm / r_0_norm < tolerance) : (r_norm < tolerance)) { converged = true; break; }
            if (iter >= max_iterations) { converged = true; break; }
            if (std::isnan(r_norm)) { converged = true; break; }

            math.vector_scaling(V[j + 1], 1.0 / alpha_j, q_k, P);
            math.vector_scaling(W[j + 1], 1.0 / alpha_j, W[j + 1], P);

            iter++;
        }

        if (j == num_vectors) j--;

        for (int k = j; k >= 0; k--)
        {
            gamma_k = gamma[k];
            for (int i = j; i > k; i--) gamma_k -= H[k][i] * c_gmres[i];
            c_gmres[k] = gamma_k / H[k][k];
        }

        // Sum Arnoldi vectors (domain.tpp:901-907)
        {
            std::vector<const double *> zp(j + 1);
            for (int i = 0; i < j + 1; i++) { zp[i] = dp(Z[i]); hcol_hst[i] = c_gmres[i]; }
            hcol.copyFrom(hcol_hst.data(), (j + 1) * sizeof(double));
            dev::check_rc(prfdd_multi_axpy_dev(dp(u_k), zp.data(), dp(hcol), 1, 1.0, j + 1, P, st()), "multi_axpy");
        }

        if (converged) break;
    }

    num_iterations = iter;
}
