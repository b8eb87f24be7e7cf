// subdomain_multi.hpp -- the multi-rank part of the PR-FDD constructor (subdomain.tpp:198-3549) and of the tree
// operator (subdomain.tpp:4566-4646): overlap rings at the ladder degrees, extended elements, superdomain and its
// AMG-composite coarsening, non-conforming region Q, interface maps, low-order FEM with hanging nodes, and the
// device-side tree exchange (NCCL send/recv + allgather instead of D2H -> MPI_Allgatherv + gslib_gs -> H2D).
//
// Where the reference gathers global data on every rank with MPI_Allgatherv (corner ids of all elements, the N = 1
// geometry and numbering, subdomain.tpp:198-262, 1632-1660) and pulls region-element data from the owners through
// gslib (644-805), this build reads the other ranks' mesh files directly: all ranks of the one 8-GPU box see the
// same directory.  The resulting tables are identical.
// Included at the end of subdomain.hpp.
#pragma once


// ---------------------------------------------------------------------------------------------
template <typename DType>
void Subdomain<DType>::build_multi_rank(std::map<int, std::unique_ptr<Domain<DType>>> &domains)
{
    using namespace prfdd_host;
    using namespace prfdd_multi;
    Domain<DType> &domain = *domains.at(poly_degree[0]);
    const std::string directory = domain.directory;
    const int num_vertices = (dim == 2) ? 4 : 8;
    const int num_edges = (dim == 2) ? 4 : 12;
    const int num_faces = (dim == 2) ? 0 : 6;
    const int N0 = poly_degree[0];
    auto npe_of = [&](int deg) { int v = 1; for (int d = 0; d < dim; d++) v *= (deg + 1); return v; };

    // elements per rank (tpp:216-223, 264-280)
    proc_count.assign(num_procs, 0);
    proc_offset.assign(num_procs, 0);
    {
        std::vector<long long> mine(1, domain.num_local_elements), all(num_procs);
        comm_world.allgather_host(mine.data(), all.data(), sizeof(long long));
        for (int p = 0; p < num_procs; p++) proc_count[p] = (int)all[p];
        for (int p = 1; p < num_procs; p++) proc_offset[p] = proc_offset[p - 1] + proc_count[p - 1];
    }
    const int T = proc_offset[num_procs - 1] + proc_count[num_procs - 1];
    num_total_elements = T;
    std::vector<std::pair<int, int>> partition(T);
    for (int p = 0; p < num_procs; p++)
        for (int e = 0; e < proc_count[p]; e++) partition[proc_offset[p] + e] = {p, e};

    FileCache<double> fd(directory);
    FileCache<long long> fl(directory);

    // corner ids of every element (tpp:225-262)
    std::vector<long long> geometry_mesh((size_t)T * num_vertices);
    {
        const std::vector<int> cidx = corner_indices(dim, N0 + 1);
        const int npe = npe_of(N0);
        for (int p = 0; p < num_procs; p++)
        {
            const std::vector<long long> &g = fl.get("glo_num", p, N0, (size_t)proc_count[p] * npe);
            for (int e = 0; e < proc_count[p]; e++)
                for (int v = 0; v < num_vertices; v++) geometry_mesh[(size_t)(proc_offset[p] + e) * num_vertices + v] = g[(size_t)e * npe + cidx[v]];
        }
        fl.clear();
    }

    // mesh connectivity (tpp:282-430)
    std::vector<std::vector<std::vector<int>>> vert_conn(T, std::vector<std::vector<int>>(num_vertices));
    std::vector<std::vector<std::vector<int>>> edge_conn(T, std::vector<std::vector<int>>(num_edges));
    std::vector<std::vector<std::vector<int>>> face_conn(T, std::vector<std::vector<int>>(num_faces));
    {
        std::map<long long, std::vector<int>> vertices;
        for (int e = 0; e < T; e++)
            for (int v = 0; v < num_vertices; v++) vertices[geometry_mesh[(size_t)e * num_vertices + v]].push_back(e);
        for (int e = 0; e < T; e++)
            for (int v = 0; v < num_vertices; v++)
                for (int x : vertices[geometry_mesh[(size_t)e * num_vertices + v]])
                    if (x != e) vert_conn[e][v].push_back(x);
        std::map<std::pair<long long, long long>, std::vector<int>> edges;
        auto ekey = [&](int e, int eid) {
            const int *pr = (dim == 2) ? EDGE_PAIRS_2D[eid] : EDGE_PAIRS_3D[eid];
            long long a = geometry_mesh[(size_t)e * num_vertices + pr[0]], b = geometry_mesh[(size_t)e * num_vertices + pr[1]];
            return std::make_pair(std::min(a, b), std::max(a, b));
        };
        for (int e = 0; e < T; e++)
            for (int eid = 0; eid < num_edges; eid++) edges[ekey(e, eid)].push_back(e);
        for (int e = 0; e < T; e++)
            for (int eid = 0; eid < num_edges; eid++)
                for (int x : edges[ekey(e, eid)])
                    if (x != e) edge_conn[e][eid].push_back(x);
        if (dim == 3)
        {
            std::map<std::array<long long, 4>, std::vector<int>> faces;
            auto fkey = [&](int e, int fid) {
                std::array<long long, 4> k;
                for (int q = 0; q < 4; q++) k[q] = geometry_mesh[(size_t)e * num_vertices + FACE_QUADS[fid][q]];
                std::sort(k.begin(), k.end());
                return k;
            };
            for (int e = 0; e < T; e++)
                for (int fid = 0; fid < 6; fid++) faces[fkey(e, fid)].push_back(e);
            for (int e = 0; e < T; e++)
                for (int fid = 0; fid < 6; fid++)
                    for (int x : faces[fkey(e, fid)])
                        if (x != e) face_conn[e][fid].push_back(x);
        }
    }
    // element adjacency incl. self: the "expander" (tpp:432-453)
    std::vector<std::vector<int>> adj(T);
    for (int e = 0; e < T; e++)
    {
        std::set<int> s;
        s.insert(e);
        for (auto &c : vert_conn[e]) s.insert(c.begin(), c.end());
        for (auto &c : edge_conn[e]) s.insert(c.begin(), c.end());
        for (auto &c : face_conn[e]) s.insert(c.begin(), c.end());
        adj[e].assign(s.begin(), s.end());
    }
    auto expand = [&](const std::vector<char> &in) {
        std::vector<char> out(T, 0);
        for (int e = 0; e < T; e++)
            if (in[e])
                for (int x : adj[e]) out[x] = 1;
        return out;
    };

    // ---- computational regions (tpp:455-553) ----------------------------------------------------
    std::vector<int> sub_ids, sub_deg, sup_ids;
    std::vector<char> marked(T, 0), reach(T, 0);
    const int off = proc_offset[proc_id], nloc = domain.num_local_elements;
    for (int e = 0; e < nloc; e++) { sub_ids.push_back(off + e); sub_deg.push_back(poly_degree[0]); marked[off + e] = 1; reach[off + e] = 1; }
    {
        int overlap = subdomain_overlap;
        for (int l = 0; l < num_levels; l++)
        {
            for (int nu = 0; nu < overlap; nu++) reach = expand(reach);
            for (int e = 0; e < T; e++)
                if (reach[e] && !marked[e]) { marked[e] = 1; sub_ids.push_back(e); sub_deg.push_back(poly_degree[l]); }
            if (overlap == 0) overlap = 1;
        }
    }
    num_subdomain_elems = (int)sub_ids.size();
    {
        std::vector<char> reach1 = expand(reach);
        for (int e = 0; e < T; e++)
            if (!marked[e])
            {
                if (reach1[e]) { sub_ids.push_back(e); sub_deg.push_back(poly_degree[num_levels - 1]); }
                sup_ids.push_back(e);
            }
    }
    num_subdomain_extended_elems = (int)sub_ids.size();
    num_superdomain_elems = (int)sup_ids.size();
    {
        std::vector<char> not_marked(T);
        for (int e = 0; e < T; e++) not_marked[e] = !marked[e];
        std::vector<char> near_sup = expand(not_marked);
        for (int e = 0; e < T; e++)
            if (marked[e] && near_sup[e]) sup_ids.push_back(e);
    }
    num_superdomain_extended_elems = (int)sup_ids.size();
    std::vector<int> subdomain_partition(T, 0);
    for (size_t k = 0; k < sub_ids.size(); k++) subdomain_partition[sub_ids[k]] = (int)k + 1;

    // pull region-element data from the owners' files (tpp:644-805)
    auto make_region = [&](const std::vector<int> &ids, const std::vector<int> *degs, std::vector<Element<DType>> &region) {
        region.clear();
        region.reserve(ids.size());
        int offs = 0;
        for (size_t k = 0; k < ids.size(); k++)
        {
            const int deg = degs ? (*degs)[k] : poly_degree[num_levels - 1];
            Element<DType> el(ids[k], dim, deg);
            el.offset = offs;
            offs += el.num_points;
            const int owner = partition[ids[k]].first, le = partition[ids[k]].second;
            const size_t cnt = (size_t)proc_count[owner] * el.num_points, src = (size_t)le * el.num_points;
            auto cpd = [&](const char *name, std::vector<DType> &dst) {
                const std::vector<double> &a = fd.get(name, owner, deg, cnt);
                for (int v = 0; v < el.num_points; v++) dst[v] = a[src + v];
            };
            cpd("x", el.x);
            if (dim >= 2) cpd("y", el.y);
            if (dim >= 3) cpd("z", el.z);
            cpd("p_mask", el.dirichlet_mask);
            for (int g = 0; g < NUM_GEOM_FACTS; g++)
            {
                char nm[16];
                snprintf(nm, sizeof(nm), "g_%d", g + 1);
                cpd(nm, el.geom_fact[g]);
            }
            const std::vector<long long> &gl = fl.get("glo_num", owner, deg, cnt);
            for (int v = 0; v < el.num_points; v++) { el.glo_num[v] = gl[src + v]; el.loc_num[v] = el.offset + v; el.dof_num[v] = 0; }
            region.push_back(std::move(el));
        }
    };
    make_region(sub_ids, &sub_deg, subdomain_region);
    make_region(sup_ids, nullptr, superdomain_region);
    num_subdomain_points = 0;
    num_subdomain_extended_points = 0;
    for (int e = 0; e < num_subdomain_extended_elems; e++)
    {
        if (e < num_subdomain_elems) num_subdomain_points += subdomain_region[e].num_points;
        num_subdomain_extended_points += subdomain_region[e].num_points;
    }
    for (int e = 0; e < num_subdomain_elems; e++) elements.push_back(subdomain_region[e]);
    for (int e = 0; e < num_superdomain_elems; e++) elements.push_back(superdomain_region[e]);

    // device copies of the region's geometric factors (tpp:667-699)
    {
        std::vector<DType> w(num_subdomain_extended_points);
        for (int g = 0; g < NUM_GEOM_FACTS; g++)
        {
            for (auto &el : subdomain_region) std::copy(el.geom_fact[g].begin(), el.geom_fact[g].end(), w.begin() + el.offset);
            subdomain_operator.geom_fact[g] = device.malloc<DType>(std::max(num_subdomain_extended_points, 1));
            subdomain_operator.geom_fact[g].copyFrom(w.data(), w.size() * sizeof(DType));
        }
    }

    // ---- interface nodes (tpp:810-843) -------------------------------------------------------------
    std::unordered_set<long long> subdomain_glo_num, interface_glo_num;
    for (int e = 0; e < num_subdomain_elems; e++)
    {
        auto &el = subdomain_region[e];
        if (el.poly_degree == 1)
            for (int v = 0; v < el.num_points; v++)
                if (el.dirichlet_mask[v] > 0.0) subdomain_glo_num.insert(el.glo_num[v]);
    }
    for (int e = 0; e < num_superdomain_elems; e++)
        for (long long g : superdomain_region[e].glo_num)
            if (subdomain_glo_num.count(g)) interface_glo_num.insert(g);
    for (auto *region : {&subdomain_region, &superdomain_region})
        for (auto &el : *region)
            for (int v = 0; v < el.num_points; v++)
                if (interface_glo_num.count(el.glo_num[v])) el.dof_num[v] = el.glo_num[v];

    // ---- connectivity of the regions (tpp:845-878) --------------------------------------------------
    for (auto *region : {&subdomain_region, &superdomain_region})
    {
        std::vector<int> mapping(T, 0);
        for (size_t k = 0; k < region->size(); k++) mapping[(*region)[k].id] = (int)k + 1;
        for (auto &el : *region)
        {
            for (int v = 0; v < num_vertices; v++)
                for (int x : vert_conn[el.id][v])
                    if (mapping[x] > 0) el.vert_conn[v].insert(mapping[x] - 1);
            for (int q = 0; q < num_edges; q++)
                for (int x : edge_conn[el.id][q])
                    if (mapping[x] > 0) el.edge_conn[q].insert(mapping[x] - 1);
            for (int q = 0; q < num_faces; q++)
                for (int x : face_conn[el.id][q])
                    if (mapping[x] > 0) el.face_conn[q].insert(mapping[x] - 1);
        }
    }

    // ---- global numbering (tpp:920-1176) -------------------------------------------------------------
    {
        std::unordered_map<int, long long> global_offset;
        global_offset[poly_degree[0]] = 0;
        for (int l = 1; l < num_levels; l++) global_offset[poly_degree[l]] = global_offset[poly_degree[l - 1]] + (long long)T * npe_of(poly_degree[l - 1]);
        for (auto &el : subdomain_region)
        {
            const std::vector<int> cidx = corner_indices(dim, el.n_x);
            std::vector<long long> corners;
            for (int c : cidx) corners.push_back(el.glo_num[c]);
            for (auto &g : el.glo_num) g += global_offset[el.poly_degree];
            for (size_t c = 0; c < cidx.size(); c++) el.glo_num[cidx[c]] = corners[c];
        }
        // zero the interiors of edges / faces shared with a lower-degree element (tpp:969-1098)
        for (auto &el : subdomain_region)
        {
            const int n = el.n_x;
            for (int q = 0; q < num_edges; q++)
            {
                bool lower = false;
                for (int j : el.edge_conn[q]) lower = lower || (subdomain_region[j].poly_degree < el.poly_degree);
                if (!lower) continue;
                const std::vector<int> ep = edge_points(dim, n, q);
                for (int k = 1; k < n - 1; k++) el.glo_num[ep[k]] = 0;
            }
            for (int q = 0; q < num_faces; q++)
            {
                bool lower = false;
                for (int j : el.face_conn[q]) lower = lower || (subdomain_region[j].poly_degree < el.poly_degree);
                if (!lower) continue;
                const std::vector<int> fp = face_points(n, q);
                for (int b = 1; b < n - 1; b++)
                    for (int a = 1; a < n - 1; a++) el.glo_num[fp[a + b * n]] = 0;
            }
        }
        // interface nodes second to last, extended nodes last (tpp:1100-1149)
        auto max_glo = [](std::vector<Element<DType>> &r) { long long m = 0; for (auto &el : r) for (auto g : el.glo_num) m = std::max(m, g); return m; };
        long long mx = max_glo(subdomain_region);
        for (auto &el : subdomain_region)
            if (el.poly_degree == 1)
                for (int v = 0; v < el.num_points; v++)
                    if (el.dof_num[v] > 0) el.glo_num[v] += mx;
        mx = max_glo(subdomain_region);
        for (int e = num_subdomain_elems; e < (int)subdomain_region.size(); e++)
        {
            auto &el = subdomain_region[e];
            for (int v = 0; v < el.num_points; v++)
                if (el.dirichlet_mask[v] > 0.0 && el.dof_num[v] == 0) el.glo_num[v] += mx;
        }
        mx = max_glo(superdomain_region);
        for (auto &el : superdomain_region)
            for (int v = 0; v < el.num_points; v++)
                if (el.dirichlet_mask[v] > 0.0 && el.dof_num[v] == 0) el.glo_num[v] += mx;
        mx = max_glo(superdomain_region);
        for (int e = num_superdomain_elems; e < (int)superdomain_region.size(); e++)
        {
            auto &el = superdomain_region[e];
            for (int v = 0; v < el.num_points; v++)
                if (el.dirichlet_mask[v] > 0.0 && el.dof_num[v] == 0) el.glo_num[v] += mx;
        }
        for (auto *region : {&subdomain_region, &superdomain_region})
        {
            int np = 0;
            for (auto &el : *region) np += el.num_points;
            std::vector<DType> w(np);
            for (auto &el : *region)
                for (int v = 0; v < el.num_points; v++) w[el.offset + v] = (DType)(el.glo_num[v]);
            ranking(w, np);
            for (auto &el : *region)
                for (int v = 0; v < el.num_points; v++) el.glo_num[v] = (long long)(w[el.offset + v]);
            for (auto &el : *region)
                for (int v = 0; v < el.num_points; v++) w[el.offset + v] = (DType)(el.glo_num[v]) * el.dirichlet_mask[v];
            ranking(w, np);
            for (auto &el : *region)
                for (int v = 0; v < el.num_points; v++) el.dof_num[v] = (long long)(w[el.offset + v]);
        }
    }

    // ---- region Q with interpolation rows on non-conforming edges / faces (tpp:1496-1585) ------------
    setup_mark("regions, rings, tree ids, renumbering");
    build_region_Q(subdomain_region, subdomain_operator.Q);
    build_region_Q(superdomain_region, superdomain_operator.Q);
    subdomain_operator.Q.transpose(subdomain_operator.Qt);
    superdomain_operator.Q.transpose(superdomain_operator.Qt);

    // ---- subdomain stiffness operator (tpp:1587-1630) -------------------------------------------------
    subdomain_operator.num_dofs = 0;
    for (int e = 0; e < num_subdomain_elems; e++)
        subdomain_operator.num_dofs = std::max(subdomain_operator.num_dofs, (int)(*std::max_element(subdomain_region[e].dof_num.begin(), subdomain_region[e].dof_num.end())));
    subdomain_operator.num_points = subdomain_operator.Q.num_rows;
    subdomain_operator.num_extended_dofs = subdomain_operator.Q.num_cols;
    {
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

    // ---- global N = 1 problem and its AMG hierarchy (tpp:1632-1858) ----------------------------------
    const int Nc = poly_degree[num_levels - 1];
    std::vector<DType> geom_fact_coarse[NUM_GEOM_FACTS];
    dof_num_coarse.assign((size_t)T * num_vertices, 0);
    std::vector<long long> glo_num_coarse((size_t)T * num_vertices, 0);
    for (int g = 0; g < NUM_GEOM_FACTS; g++) geom_fact_coarse[g].resize((size_t)T * num_vertices);
    for (int p = 0; p < num_procs; p++)
    {
        const size_t cnt = (size_t)proc_count[p] * num_vertices, o = (size_t)proc_offset[p] * num_vertices;
        for (int g = 0; g < NUM_GEOM_FACTS; g++)
        {
            char nm[16];
            snprintf(nm, sizeof(nm), "g_%d", g + 1);
            const std::vector<double> &a = fd.get(nm, p, Nc, cnt);
            std::copy(a.begin(), a.end(), geom_fact_coarse[g].begin() + o);
        }
        const std::vector<double> &m = fd.get("p_mask", p, Nc, cnt);
        const std::vector<long long> &gl = fl.get("glo_num", p, Nc, cnt);
        for (size_t i = 0; i < cnt; i++) glo_num_coarse[o + i] = (m[i] > 0.0) ? gl[i] : 0;
    }
    int num_coarse_dofs = 0;
    {
        // integer dense ranking (tpp:1666-1704)
        std::vector<long long> sorted(glo_num_coarse);
        std::sort(sorted.begin(), sorted.end());
        sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
        const long long base = (sorted[0] == 0) ? 0 : 1;
        for (size_t i = 0; i < glo_num_coarse.size(); i++)
            dof_num_coarse[i] = (long long)(std::lower_bound(sorted.begin(), sorted.end(), glo_num_coarse[i]) - sorted.begin()) + base;
        for (auto d : dof_num_coarse) num_coarse_dofs = std::max(num_coarse_dofs, (int)d);
    }
    Qt_coarse.initialize(num_coarse_dofs, T * num_vertices);
    for (int e = 0; e < T; e++)
        for (int v = 0; v < num_vertices; v++)
            if (dof_num_coarse[(size_t)e * num_vertices + v] > 0) Qt_coarse.add_entry((int)dof_num_coarse[(size_t)e * num_vertices + v] - 1, e * num_vertices + v, 1.0);
    Qt_coarse.assemble();

    amg::Hierarchy amg_coarse;
    amg::HostCSR A0;
    {
        std::vector<std::tuple<int, int, double>> coo;
        std::vector<DType> Ae(num_vertices * num_vertices);
        for (int e = 0; e < T; e++)
        {
            const DType *gp[NUM_GEOM_FACTS];
            for (int g = 0; g < NUM_GEOM_FACTS; g++) gp[g] = geom_fact_coarse[g].data() + (size_t)e * num_vertices;
            q1_element_matrix(gp, Ae);
            for (int i = 0; i < num_vertices; i++)
                for (int j = 0; j < num_vertices; j++)
                {
                    const int row = (int)dof_num_coarse[(size_t)e * num_vertices + i] - 1, col = (int)dof_num_coarse[(size_t)e * num_vertices + j] - 1;
                    const DType val = Ae[i * num_vertices + j];
                    if (row >= 0 && col >= 0 && std::abs(val) > epsilon) coo.emplace_back(row, col, val);
                }
        }
        A0 = csr_from_coo(num_coarse_dofs, num_coarse_dofs, coo);
    }
    // BoomerAMG #1 stand-in: coarsen to a single dof (tpp:1851-1858)
    setup_mark("region Q, operators, global N=1 matrix");
    amg_coarse.coarsening = opt.amg_coarsening;
    amg_coarse.setup(A0, 1, /*max_coarse=*/1, 0.25, 4, 25, /*on_device=*/false);
    setup_mark("AMG #1 (global N=1 matrix)");

    build_superdomain(amg_coarse, sub_ids, sup_ids, interface_glo_num, glo_num_coarse, num_coarse_dofs);
    setup_mark("composite superdomain grid");
    build_interface(sub_ids, sup_ids, subdomain_partition, (int)interface_glo_num.size());
    setup_mark("interface maps, weights");

    // ---- low-order preconditioner (tpp:2749-3549) ------------------------------------------------------
    rstdout("Assembling subdomain low-order preconditioner\n");
    if (use_preconditioner)
    {
        assemble_low_order_fem();
        setup_mark("low-order FEM assembly");
        amg_fem.coarsening = opt.amg_coarsening;
        amg_fem.fp32 = opt.amg_precision == 1;
        amg_fem.setup(A_fem_hst, cheby_order);
        setup_mark("AMG #2 (hierarchy, upload, collapsed coarse levels)");
    }

    setup_tree_exchange();
    setup_mark("tree exchange lists");
}

// Q1 SEM element matrix D^T G D of an N = 1 element (tpp:1715-1826, 3040-3124); Ae row-major nv x nv
template <typename DType>
void Subdomain<DType>::q1_element_matrix(const DType *const g[NUM_GEOM_FACTS], std::vector<DType> &Ae)
{
    const int dimn = prfdd_host::dim;
    const int nv = (dimn == 2) ? 4 : 8;
    const std::vector<DType> &d2 = D_hat[num_levels - 1].first;
    std::vector<DType> Dm[3];
    for (int c = 0; c < dimn; c++) Dm[c].assign(nv * nv, 0.0);
    if (dimn == 2)
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
    Ae.assign(nv * nv, 0.0);
    for (int i = 0; i < nv; i++)
        for (int j = 0; j < nv; j++)
        {
            DType val = 0.0;
            for (int k = 0; k < nv; k++)
                for (int a = 0; a < dimn; a++)
                {
                    DType gd = 0.0;
                    for (int b = 0; b < dimn; b++) gd += g[dimn == 2 ? gi2[a][b] : gi3[a][b]][k] * Dm[b][k * nv + j];
                    val += Dm[a][k * nv + i] * gd;
                }
            Ae[i * nv + j] = val;
        }
}

// local indices on elem_j of the edge (kind 0) / face (kind 1) `idx` of elem_i, identified through the corner ids and
// assumed identically oriented (subdomain.tpp:1179-1494)
template <typename DType>
std::pair<std::vector<int>, std::vector<int>> Subdomain<DType>::matching(const Element<DType> &ei, const Element<DType> &ej, int kind, int idx)
{
    using namespace prfdd_multi;
    const int dimn = prfdd_host::dim;
    const int ni = ei.n_x, nj = ej.n_x;
    if (kind == 0)
    {
        std::vector<int> pi = edge_points(dimn, ni, idx);
        const long long a = ei.glo_num[pi[0]], b = ei.glo_num[pi[ni - 1]];
        for (int q = 0; q < (dimn == 2 ? 4 : 12); q++)
        {
            std::vector<int> pj = edge_points(dimn, nj, q);
            const long long c = ej.glo_num[pj[0]], d = ej.glo_num[pj[nj - 1]];
            if ((c == a || c == b) && (d == a || d == b)) return {pi, pj};
        }
        throw std::runtime_error("Subdomain: matching_edge found no common edge");
    }
    std::vector<int> pi = face_points(ni, idx);
    const long long cs[4] = {ei.glo_num[pi[0]], ei.glo_num[pi[ni - 1]], ei.glo_num[pi[(ni - 1) * ni]], ei.glo_num[pi[ni * ni - 1]]};
    auto in = [&](long long v) { return v == cs[0] || v == cs[1] || v == cs[2] || v == cs[3]; };
    for (int q = 0; q < 6; q++)
    {
        std::vector<int> pj = face_points(nj, q);
        if (in(ej.glo_num[pj[0]]) && in(ej.glo_num[pj[nj - 1]]) && in(ej.glo_num[pj[(nj - 1) * nj]]) && in(ej.glo_num[pj[nj * nj - 1]])) return {pi, pj};
    }
    throw std::runtime_error("Subdomain: matching_face found no common face");
}

template <typename DType>
int Subdomain<DType>::min_degree_edge_neighbor(const std::vector<Element<DType>> &region, const Element<DType> &el, int eid)
{
    int e_j = -1, N_j = el.poly_degree;
    for (int e : el.edge_conn[eid])
        if (region[e].poly_degree < N_j) { e_j = e; N_j = region[e].poly_degree; }
    return e_j;
}

template <typename DType>
void Subdomain<DType>::build_region_Q(std::vector<Element<DType>> &region, CSR_Matrix<DType> &Q)
{
    const int dimn = prfdd_host::dim;
    const int num_edges = (dimn == 2) ? 4 : 12, num_faces = (dimn == 2) ? 0 : 6;
    int num_points = region.empty() ? 0 : region.back().offset + region.back().num_points;
    int ndofs = 0;
    for (auto &el : region)
        for (auto d : el.dof_num) ndofs = std::max(ndofs, (int)d);
    Q.initialize(num_points, ndofs);
    for (auto &ei : region)
    {
        const int Ni = ei.poly_degree, ni = Ni + 1;
        for (int v = 0; v < ei.num_points; v++)
            if (ei.dof_num[v] > 0) Q.add_entry(ei.loc_num[v], (int)ei.dof_num[v] - 1, 1.0);
        for (int q = 0; q < num_edges; q++)
        {
            const int e_j = min_degree_edge_neighbor(region, ei, q);
            if (e_j < 0) continue;
            const Element<DType> &ej = region[e_j];
            const int Nj = ej.poly_degree, nj = Nj + 1;
            auto m = matching(ei, ej, 0, q);
            const std::vector<DType> &J = J_cf[std::pair<int, int>(Nj, Ni)].first;
            for (int i = 1; i < ni - 1; i++)
                for (int j = 0; j < nj; j++)
                    if (ej.dof_num[m.second[j]] > 0) Q.add_entry(ei.loc_num[m.first[i]], (int)ej.dof_num[m.second[j]] - 1, J[i * nj + j]);
        }
        for (int q = 0; q < num_faces; q++)
            for (int e_j : ei.face_conn[q])
            {
                const Element<DType> &ej = region[e_j];
                const int Nj = ej.poly_degree, nj = Nj + 1;
                if (Ni <= Nj) continue;
                auto m = matching(ei, ej, 1, q);
                const std::vector<DType> &J = J_cf[std::pair<int, int>(Nj, Ni)].first;
                for (int j = 1; j < ni - 1; j++)
                    for (int i = 1; i < ni - 1; i++)
                        for (int qq = 0; qq < nj; qq++)
                            for (int pp = 0; pp < nj; pp++)
                                if (ej.dof_num[m.second[pp + qq * nj]] > 0)
                                    Q.add_entry(ei.loc_num[m.first[i + j * ni]], (int)ej.dof_num[m.second[pp + qq * nj]] - 1, J[i * nj + pp] * J[j * nj + qq]);
            }
    }
    Q.assemble();
}

// ---- superdomain composite grid through the AMG hierarchy of the global N = 1 problem (tpp:1860-2579) ----
template <typename DType>
void Subdomain<DType>::build_superdomain(amg::Hierarchy &H, const std::vector<int> &sub_ids, const std::vector<int> &sup_ids,
                                         const std::unordered_set<long long> &interface_glo_num, const std::vector<long long> &glo_num_coarse, int ncd)
{
    using namespace prfdd_multi;
    using amg::HostCSR;
    const int nv = (prfdd_host::dim == 2) ? 4 : 8;
    dof_marker.assign(ncd, 0);
    for (int e = 0; e < num_subdomain_elems; e++)
        for (int v = 0; v < nv; v++)
        {
            const long long dof = dof_num_coarse[(size_t)sub_ids[e] * nv + v], glo = glo_num_coarse[(size_t)sub_ids[e] * nv + v];
            if (dof > 0) dof_marker[dof - 1] = 1;
            if (interface_glo_num.count(glo)) dof_marker[dof - 1] = 2;
        }
    for (int e = num_subdomain_elems; e < num_subdomain_extended_elems; e++)
        for (int v = 0; v < nv; v++)
        {
            const long long dof = dof_num_coarse[(size_t)sub_ids[e] * nv + v];
            if (dof > 0 && dof_marker[dof - 1] == 0) dof_marker[dof - 1] = 3;
        }
    for (int e = num_superdomain_elems; e < num_superdomain_extended_elems; e++)
        for (int v = 0; v < nv; v++)
        {
            const long long dof = dof_num_coarse[(size_t)sup_ids[e] * nv + v];
            if (dof > 0 && dof_marker[dof - 1] == 1) dof_marker[dof - 1] = 4;
        }

    const int nlev = H.num_levels();
    std::vector<int> num_nodes(nlev);
    for (int l = 0; l < nlev; l++) num_nodes[l] = H.levels[l].n;
    std::vector<std::vector<double>> D(nlev);
    for (int l = 0; l < nlev; l++) D[l].assign(num_nodes[l], 0.0);
    for (int i = 0; i < num_nodes[0]; i++)
        if (dof_marker[i] > 0) D[0][i] = 1.0;
    // C-points of level l and their coarse index (the reference's test is "P row has exactly one entry")
    auto c_rows = [&](int l, std::vector<int> &rows, std::vector<int> &cols) {
        rows.clear(); cols.clear();
        const amg::Level &L = H.levels[l];
        for (int r = 0; r < L.n; r++)
            if (L.cf[r] == 1) { rows.push_back(r); cols.push_back(L.P.col[L.P.ptr[r]]); }
    };
    int ncl = 0;
    int ov = superdomain_overlap;
    std::vector<int> rows, cols;
    for (int l = 0; l < nlev; l++)
    {
        ncl = l + 1;
        const HostCSR &A = H.levels[l].A;
        std::vector<double> w(D[l]), w2(num_nodes[l]);
        for (int nu = 0; nu < ov; nu++)
        {
            for (int r = 0; r < num_nodes[l]; r++)
            {
                double s = 0.0;
                for (int k = A.ptr[r]; k < A.ptr[r + 1]; k++) s += w[A.col[k]];
                w2[r] = s;
            }
            w = w2;
        }
        if (ov == 0) ov = 1;
        if (l == nlev - 1) std::fill(w.begin(), w.end(), 1.0);
        bool any_zero = false;
        for (int i = 0; i < num_nodes[l]; i++)
        {
            if (D[l][i] == 0.0 && w[i] > 0.0) D[l][i] = 2.0;
            if (D[l][i] == 0.0) any_zero = true;
        }
        if (!any_zero) break;
        if (l < nlev - 1)
        {
            c_rows(l, rows, cols);
            for (size_t k = 0; k < rows.size(); k++)
                if (D[l][rows[k]] > 0.0) D[l + 1][cols[k]] = 1.0;
        }
    }
    std::vector<int> num_local(ncl, 0), num_overlap(ncl, 0), num_remaining(ncl, 0);
    for (int l = 0; l < ncl; l++)
        for (int i = 0; i < num_nodes[l]; i++)
        {
            if (D[l][i] == 1.0) num_local[l]++;
            if (D[l][i] == 2.0) num_overlap[l]++;
            if (D[l][i] == 0.0) num_remaining[l]++;
        }
    std::vector<int> num_comp_overlap(num_overlap);
    num_comp_overlap[0] += num_local[0];

    std::vector<std::vector<int>> nodes_to_fine(ncl), nodes_to_dofs(ncl);
    nodes_to_fine[0].resize(num_nodes[0]);
    for (int i = 0; i < num_nodes[0]; i++) nodes_to_fine[0][i] = i;
    for (int l = 0; l < ncl - 1; l++)
    {
        nodes_to_fine[l + 1].assign(num_nodes[l + 1], 0);
        c_rows(l, rows, cols);
        for (size_t k = 0; k < rows.size(); k++) nodes_to_fine[l + 1][cols[k]] = nodes_to_fine[l][rows[k]];
    }
    for (int l = 0; l < ncl; l++) nodes_to_dofs[l].assign(num_nodes[l], -1);
    int dof_end = 0;
    for (int marker = 1; marker <= 4; marker++)
        for (int i = 0; i < num_nodes[0]; i++)
            if (dof_marker[i] == marker) nodes_to_dofs[0][i] = dof_end++;
    dof_end = num_local[0];
    for (int i = 0; i < num_nodes[0]; i++)
        if (D[0][i] == 2.0) nodes_to_dofs[0][i] = dof_end++;
    int offset = num_local[0] + num_overlap[0];
    for (int l = 0; l < ncl - 1; l++)
        for (int i = 0; i < num_nodes[l + 1]; i++)
            if (D[l + 1][i] == 2.0) nodes_to_dofs[0][nodes_to_fine[l + 1][i]] = offset++;
    for (int l = 0; l < ncl - 1; l++)
    {
        c_rows(l, rows, cols);
        for (size_t k = 0; k < rows.size(); k++) nodes_to_dofs[l + 1][cols[k]] = nodes_to_dofs[l][rows[k]];
    }
    const int num_comp_dofs = offset;

    std::vector<HostCSR> P_c(std::max(ncl - 1, 0)), R_c(std::max(ncl - 1, 0));
    for (int l = ncl - 1; l > 0; l--)
    {
        const HostCSR &Pl = H.levels[l - 1].P;
        const int nf = num_nodes[l - 1], ncn = num_nodes[l];
        std::vector<int> fine(nf, -1), coarse(ncn, -1);
        int de;
        if (l - 1 == 0)
        {
            de = 0;
            for (int marker = 1; marker <= 4; marker++)
                for (int i = 0; i < nf; i++)
                    if (dof_marker[i] == marker) fine[i] = de++;
            for (int i = 0; i < nf; i++)
                if (D[0][i] == 2.0) fine[i] = de++;
            de = num_local[0] + num_overlap[0];
            for (int i = 0; i < nf; i++)
                if (D[0][i] == 0.0) fine[i] = de++;
        }
        else
        {
            de = 0;
            for (int i = 0; i < nf; i++)
                if (D[l - 1][i] == 2.0) fine[i] = de++;
            de = num_overlap[l - 1];
            for (int i = 0; i < nf; i++)
                if (D[l - 1][i] == 0.0) fine[i] = de++;
        }
        const int bound = (l - 1 == 0) ? num_local[0] + num_overlap[0] : num_overlap[l - 1];
        de = bound;
        for (int i = 0; i < ncn; i++)
            if (D[l][i] == 2.0 || D[l][i] == 0.0) coarse[i] = de++;
        c_rows(l - 1, rows, cols);
        for (size_t k = 0; k < rows.size(); k++)
            if (fine[rows[k]] < bound) coarse[cols[k]] = fine[rows[k]];
        const int num_fine = num_overlap[l - 1] + ((l - 1 == 0) ? num_local[0] : 0);
        int nr_, nc_;
        if (l - 1 == 0) { nr_ = num_nodes[0]; nc_ = num_local[0] + num_overlap[0] + num_overlap[l] + num_remaining[l]; }
        else { nr_ = num_overlap[l - 1] + num_remaining[l - 1]; nc_ = num_overlap[l - 1] + num_overlap[l] + num_remaining[l]; }
        std::vector<std::tuple<int, int, double>> coo;
        for (int row = 0; row < nf; row++)
        {
            const int fr = fine[row];
            if (fr < 0) continue;
            if (fr < num_fine) coo.emplace_back(fr, fr, 1.0);
            else
                for (int k = Pl.ptr[row]; k < Pl.ptr[row + 1]; k++)
                    if (coarse[Pl.col[k]] >= 0) coo.emplace_back(fr, coarse[Pl.col[k]], Pl.val[k]);
        }
        P_c[l - 1] = csr_from_coo(nr_, nc_, coo);
        coo.clear();
        int cnt = 0;
        for (int i = 0; i < nf; i++)
            if (fine[i] >= 0) coo.emplace_back(cnt++, fine[i], 1.0);
        R_c[l - 1] = csr_from_coo(cnt, cnt, coo);
    }

    HostCSR Pfull;
    if (ncl > 1)
    {
        for (int l = ncl - 2; l > 0; l--)
        {
            const HostCSR &Pc = P_c[l - 1];
            const int nco = num_comp_overlap[l - 1];
            std::vector<std::tuple<int, int, double>> c21, c22;
            for (int i = nco; i < Pc.num_rows; i++)
                for (int k = Pc.ptr[i]; k < Pc.ptr[i + 1]; k++)
                {
                    if (Pc.col[k] < nco) c21.emplace_back(i - nco, Pc.col[k], Pc.val[k]);
                    else c22.emplace_back(i - nco, Pc.col[k] - nco, Pc.val[k]);
                }
            HostCSR P21 = csr_from_coo(Pc.num_rows - nco, nco, c21);
            HostCSR P22 = csr_from_coo(Pc.num_rows - nco, Pc.num_cols - nco, c22);
            HostCSR RlPl = amg::spgemm(R_c[l], P_c[l]);
            HostCSR P22n = amg::spgemm(P22, RlPl);
            std::vector<std::tuple<int, int, double>> coo;
            for (int r = 0; r < nco; r++) coo.emplace_back(r, r, 1.0);
            for (int i = 0; i < P21.num_rows; i++)
                for (int k = P21.ptr[i]; k < P21.ptr[i + 1]; k++) coo.emplace_back(i + nco, P21.col[k], P21.val[k]);
            for (int i = 0; i < P22n.num_rows; i++)
                for (int k = P22n.ptr[i]; k < P22n.ptr[i + 1]; k++) coo.emplace_back(i + nco, P22n.col[k] + nco, P22n.val[k]);
            P_c[l - 1] = csr_from_coo(Pc.num_rows, nco + P22n.num_cols, coo);
        }
        Pfull = amg::spgemm(R_c[0], P_c[0]);
    }
    else
    {
        std::vector<std::tuple<int, int, double>> coo;
        for (int i = 0; i < num_comp_dofs; i++) coo.emplace_back(i, nodes_to_dofs[0][i], 1.0);
        Pfull = csr_from_coo(num_comp_dofs, num_comp_dofs, coo);
    }
    HostCSR Pt_full = amg::transpose(Pfull);
    HostCSR PtAP = amg::spgemm(amg::spgemm(Pt_full, H.levels[0].A), Pfull);

    int marker_count[5] = {0, 0, 0, 0, 0}, marker_offset[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < num_nodes[0]; i++)
        if (dof_marker[i] >= 1 && dof_marker[i] <= 4) marker_count[dof_marker[i] - 1]++;
    for (int m = 1; m < 5; m++) marker_offset[m] = marker_offset[m - 1] + marker_count[m - 1];
    const int nrows = PtAP.num_rows;
    marker_count[4] = nrows - marker_offset[4];
    std::vector<int> R_sup(nrows, -1);
    int dof = 0;
    for (int i = marker_offset[1]; i < marker_offset[3]; i++) R_sup[i] = dof++;
    for (int i = marker_offset[4]; i < nrows; i++) R_sup[i] = dof++;
    for (int i = marker_offset[3]; i < marker_offset[4]; i++) R_sup[i] = dof++;
    const int ncols = marker_count[1] + marker_count[2] + marker_count[3] + marker_count[4];

    superdomain_operator.A.initialize(ncols, ncols);
    for (int i = 0; i < nrows; i++)
        for (int k = PtAP.ptr[i]; k < PtAP.ptr[i + 1]; k++)
            if (R_sup[i] >= 0 && R_sup[PtAP.col[k]] >= 0) superdomain_operator.A.add_entry(R_sup[i], R_sup[PtAP.col[k]], PtAP.val[k]);
    superdomain_operator.A.assemble();
    superdomain_operator.Pt.initialize(dof, Pfull.num_rows);
    for (int r = 0; r < Pfull.num_rows; r++)
        for (int k = Pfull.ptr[r]; k < Pfull.ptr[r + 1]; k++)
            if (R_sup[Pfull.col[k]] >= 0) superdomain_operator.Pt.add_entry(R_sup[Pfull.col[k]], r, Pfull.val[k]);
    superdomain_operator.Pt.assemble();

    // dof_sup: composite numbering of the coarse dofs seen from the superdomain (tpp:2481-2529)
    dof_sup.assign(nodes_to_dofs[0].begin(), nodes_to_dofs[0].end());
    int dof_max = 0;
    for (int v : nodes_to_dofs[0]) dof_max = std::max(dof_max, v);
    for (int i = 0; i < num_nodes[0]; i++)
    {
        if (dof_marker[i] == 1) dof_sup[i] = -1;
        if (dof_marker[i] == 4) dof_sup[i] += dof_max;
    }
    {
        std::vector<int> sorted(dof_sup);
        std::sort(sorted.begin(), sorted.end());
        sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
        const int base = (sorted[0] == -1) ? 0 : 1;
        for (auto &v : dof_sup) v = (int)(std::lower_bound(sorted.begin(), sorted.end(), v) - sorted.begin()) + base;
    }
    superdomain_operator.num_dofs = ncols - marker_count[3];
    superdomain_operator.num_extended_dofs = dof;
    superdomain_operator.num_points = superdomain_operator.Q.num_rows;
    num_comp_levels = ncl;
}

// ---- interface operator and weights (tpp:2581-2747) --------------------------------------------------------
template <typename DType>
void Subdomain<DType>::build_interface(const std::vector<int> &sub_ids, const std::vector<int> &sup_ids, const std::vector<int> &subdomain_partition, int n_interface)
{
    using namespace prfdd_host;
    const int nv = (dim == 2) ? 4 : 8;
    num_interface_dofs = n_interface;
    num_dofs = subdomain_operator.num_dofs + superdomain_operator.num_dofs - num_interface_dofs;
    const int shift = subdomain_operator.num_dofs - num_interface_dofs;
    const int ne = subdomain_operator.num_extended_dofs, ns = superdomain_operator.num_extended_dofs;
    std::unordered_map<long long, long long> sub_map, sup_map;
    for (int e = 0; e < num_subdomain_elems; e++)
        for (auto d : subdomain_region[e].dof_num)
            if (d > 0) sub_map[d] = d;
    for (int e = num_subdomain_elems; e < num_subdomain_extended_elems; e++)
    {
        auto &el = subdomain_region[e];
        for (int v = 0; v < el.num_points; v++)
        {
            const int dof = (int)dof_num_coarse[(size_t)el.id * nv + v];
            if (dof > 0 && dof_sup[dof - 1] > 0) sub_map[el.dof_num[v]] = dof_sup[dof - 1] + shift;
        }
    }
    for (int e = 0; e < num_superdomain_elems; e++)
    {
        auto &el = superdomain_region[e];
        for (int v = 0; v < el.num_points; v++)
        {
            const int dof = (int)dof_num_coarse[(size_t)el.id * nv + v];
            if (dof > 0) sup_map[dof_sup[dof - 1]] = dof_sup[dof - 1] + shift;
        }
    }
    for (int e = num_superdomain_elems; e < num_superdomain_extended_elems; e++)
    {
        auto &el = superdomain_region[e];
        auto &sl = subdomain_region[subdomain_partition[el.id] - 1];
        for (int v = 0; v < el.num_points; v++)
        {
            const int dof = (int)dof_num_coarse[(size_t)el.id * nv + v];
            if (dof > 0 && dof_marker[dof - 1] == 4) sup_map[dof_sup[dof - 1]] = sl.dof_num[v];
        }
    }
    Q_int.initialize(ne + ns, num_dofs);
    for (int i = 0; i < ne; i++) Q_int.add_entry(i, (int)sub_map.at(i + 1) - 1, 1.0);
    for (int i = 0; i < ns; i++) Q_int.add_entry(ne + i, (int)sup_map.at(i + 1) - 1, 1.0);
    Q_int.assemble();
    Qt_int.initialize(num_dofs, ne + ns);
    for (int i = 0; i < subdomain_operator.num_dofs; i++) Qt_int.add_entry(i, i, 1.0);
    for (int i = 0; i < superdomain_operator.num_dofs - num_interface_dofs; i++) Qt_int.add_entry(subdomain_operator.num_dofs + i, ne + num_interface_dofs + i, 1.0);
    Qt_int.assemble();
    QQt_int.initialize(ne + ns, ne + ns);
    std::vector<char> seen(ne + ns, 0);
    for (int i = 0; i < subdomain_operator.num_dofs; i++) { QQt_int.add_entry(i, i, 1.0); seen[i] = 1; }
    for (int e = num_subdomain_elems; e < num_subdomain_extended_elems; e++)
    {
        auto &el = subdomain_region[e];
        for (int v = 0; v < el.num_points; v++)
            if (el.dof_num[v] > 0 && !seen[el.dof_num[v] - 1])
            {
                QQt_int.add_entry((int)el.dof_num[v] - 1, ne + dof_sup[dof_num_coarse[(size_t)el.id * nv + v] - 1] - 1, 1.0);
                seen[el.dof_num[v] - 1] = 1;
            }
    }
    for (int i = 0; i < num_interface_dofs; i++) { QQt_int.add_entry(ne + i, subdomain_operator.num_dofs - num_interface_dofs + i, 1.0); seen[ne + i] = 1; }
    for (int i = num_interface_dofs; i < superdomain_operator.num_dofs; i++) { QQt_int.add_entry(ne + i, ne + i, 1.0); seen[ne + i] = 1; }
    for (int e = num_superdomain_elems; e < num_superdomain_extended_elems; e++)
    {
        auto &el = superdomain_region[e];
        auto &sl = subdomain_region[subdomain_partition[el.id] - 1];
        for (int v = 0; v < el.num_points; v++)
            if (dof_num_coarse[(size_t)el.id * nv + v] > 0)
            {
                const int dof = dof_sup[dof_num_coarse[(size_t)el.id * nv + v] - 1];
                if (!seen[ne + dof - 1])
                {
                    QQt_int.add_entry(ne + dof - 1, (int)sl.dof_num[v] - 1, 1.0);
                    seen[ne + dof - 1] = 1;
                }
            }
    }
    QQt_int.assemble();
    interface_is_identity = false;

    // weights (tpp:2731-2747)
    norm_weight_hst.assign(ne + ns, 1.0);
    for (int i = subdomain_operator.num_dofs; i < ne; i++) norm_weight_hst[i] = 0.0;
    for (int i = 0; i < num_interface_dofs; i++) norm_weight_hst[ne + i] = 0.0;
    for (int i = superdomain_operator.num_dofs; i < ns; i++) norm_weight_hst[ne + i] = 0.0;
    norm_weight = device.malloc<DType>(std::max(ne + ns, 1));
    norm_weight.copyFrom(norm_weight_hst.data(), (ne + ns) * sizeof(DType));
    num_values = subdomain_operator.num_points + ns;
    inner_weight = device.malloc<DType>(std::max(num_values, 1));
    subdomain_operator.Q.multiply(inner_weight, norm_weight);
    inner_weight_hst.resize(num_values);
    inner_weight.copyTo(inner_weight_hst.data(), num_values * sizeof(DType));
    for (int i = 0; i < ns; i++) inner_weight_hst[subdomain_operator.num_points + i] = norm_weight_hst[ne + i];
    for (auto &w : inner_weight_hst)
        if (w > 0.0) w = 1.0;
    inner_weight.copyFrom(inner_weight_hst.data(), num_values * sizeof(DType));
}

// ---- device-side tree exchange set-up: who sends which element blocks to whom ----------------------------------
template <typename DType>
void Subdomain<DType>::setup_tree_exchange()
{
    using namespace prfdd_host;
    auto npe_of = [&](int deg) { int v = 1; for (int d = 0; d < dim; d++) v *= (deg + 1); return v; };
    std::unordered_map<int, int> level_degree;
    for (int l = 0; l < num_levels; l++) level_degree[poly_degree[l]] = l;
    // every rank publishes its region as (global element id, degree) pairs
    long long my_n = (long long)subdomain_region.size();
    long long max_n = comm_world.allreduce_max_host(my_n);
    std::vector<long long> mine(1 + 2 * max_n, 0), all((size_t)(1 + 2 * max_n) * num_procs);
    mine[0] = my_n;
    for (long long k = 0; k < my_n; k++) { mine[1 + 2 * k] = subdomain_region[k].id; mine[2 + 2 * k] = subdomain_region[k].poly_degree; }
    comm_world.allgather_host(mine.data(), all.data(), mine.size() * sizeof(long long));

    const int off = proc_offset[proc_id], nloc = proc_count[proc_id];
    std::vector<int> send_idx, recv_idx;
    tree.peers.clear(); tree.send_count.clear(); tree.recv_count.clear(); tree.send_offset.clear(); tree.recv_offset.clear();
    for (int q = 0; q < num_procs; q++)
    {
        if (q == proc_id) continue;
        const long long *rec = all.data() + (size_t)q * mine.size();
        const size_t s0 = send_idx.size(), r0 = recv_idx.size();
        // what q needs from me, in q's region order
        for (long long k = 0; k < rec[0]; k++)
        {
            const int gid = (int)rec[1 + 2 * k], deg = (int)rec[2 + 2 * k];
            if (gid < off || gid >= off + nloc) continue;
            const int l = level_degree.at(deg), n_ = npe_of(deg);
            const int base = levels[l].offset + (gid - off) * n_;
            for (int v = 0; v < n_; v++) send_idx.push_back(base + v);
        }
        // what I need from q, in my region order
        for (auto &el : subdomain_region)
        {
            if (el.id < proc_offset[q] || el.id >= proc_offset[q] + proc_count[q]) continue;
            for (int v = 0; v < el.num_points; v++) recv_idx.push_back(el.offset + v);
        }
        if (send_idx.size() == s0 && recv_idx.size() == r0) continue;
        tree.peers.push_back(q);
        tree.send_offset.push_back((int)s0); tree.send_count.push_back((int)(send_idx.size() - s0));
        tree.recv_offset.push_back((int)r0); tree.recv_count.push_back((int)(recv_idx.size() - r0));
    }
    tree.send_total = (int)send_idx.size();
    tree.recv_total = (int)recv_idx.size();
    tree.send_idx = device.malloc<int>(std::max(tree.send_total, 1));
    tree.send_idx.copyFrom(send_idx.data(), send_idx.size() * sizeof(int));
    tree.recv_idx = device.malloc<int>(std::max(tree.recv_total, 1));
    tree.recv_idx.copyFrom(recv_idx.data(), recv_idx.size() * sizeof(int));
    tree.send_buf = device.malloc<double>(std::max(tree.send_total, 1));
    tree.recv_buf = device.malloc<double>(std::max(tree.recv_total, 1));
    const int nv = (dim == 2) ? 4 : 8;
    for (int p = 0; p < num_procs; p++)
        if (proc_count[p] != proc_count[0]) throw std::runtime_error("Subdomain: the coarse allgather needs the same number of elements on every rank");
    tree.coarse_per_rank = proc_count[0] * nv;
    tree.coarse_all = device.malloc<double>((size_t)tree.coarse_per_rank * num_procs);
    tree.coarse_dofs = device.malloc<double>(std::max(Qt_coarse.num_rows, 1));
    tree.level_buf = device.malloc<double>(std::max(levels.back().offset + levels.back().num_points, 1));
}

// subdomain.tpp:4566-4646 on the device: cast, ladder restrictions, NCCL exchange of the region blocks, allgather of the
// N = 1 level, Qt_coarse and Pt products.  No host staging.
template <typename DType>
void Subdomain<DType>::tree_operator_multi(const memory &Tu, const memory &u)
{
    using namespace prfdd_host;
    double *lev = dp(tree.level_buf);
    dev::check_rc(prfdd_copy_from_domain_data(lev, dp(u), own_points, st()), "copy_from_domain_data");
    timer.start("subdomain.tree_construction.subdomain");
    for (int l = 0; l < num_levels - 1; l++)
    {
        const int n_f = levels[l].poly_degree + 1, n_c = levels[l + 1].poly_degree + 1;
        const memory &J = J_cf[std::pair<int, int>(levels[l + 1].poly_degree, levels[l].poly_degree)].second;
        dev::check_rc(prfdd_restriction(lev + levels[l + 1].offset, dp(J), lev + levels[l].offset, levels[l].num_elements, n_f, n_c, dim, st()), "restriction");
    }
    timer.stop("subdomain.tree_construction.subdomain");

    timer.start("subdomain.tree_exchange.subdomain");
    dev::check_rc(prfdd_halo_pack(dp(tree.send_buf), lev, tree.send_idx.template as<int>(), tree.send_total, st()), "tree pack");
    {
        std::vector<const double *> sp;
        std::vector<double *> rp;
        std::vector<size_t> sc, rc;
        for (size_t k = 0; k < tree.peers.size(); k++)
        {
            sp.push_back(dp(tree.send_buf) + tree.send_offset[k]); sc.push_back((size_t)tree.send_count[k]);
            rp.push_back(dp(tree.recv_buf) + tree.recv_offset[k]); rc.push_back((size_t)tree.recv_count[k]);
        }
        comm_world.sendrecv(tree.peers, sp, sc, rp, rc);
    }
    timer.stop("subdomain.tree_exchange.subdomain");
    timer.start("subdomain.tree_exchange.superdomain");
    comm_world.allgather(lev + levels[num_levels - 1].offset, dp(tree.coarse_all), (size_t)tree.coarse_per_rank);
    timer.stop("subdomain.tree_exchange.superdomain");

    // own elements at degree N are the first region slots
    dev::check_rc(prfdd_memcpy_d2d(dp(Tu), lev, (size_t)own_points * sizeof(double), st()), "tree own copy");
    dev::check_rc(prfdd_scatter_assign(dp(Tu), dp(tree.recv_buf), tree.recv_idx.template as<int>(), tree.recv_total, st()), "tree unpack");

    timer.start("subdomain.tree_construction.assemble_coarse");
    Qt_coarse.multiply(tree.coarse_dofs, tree.coarse_all);
    timer.stop("subdomain.tree_construction.assemble_coarse");
    timer.start("subdomain.tree_construction.superdomain");
    if (superdomain_operator.num_extended_dofs > 0)
    {
        memory Tu_sup = Tu.slice(subdomain_operator.num_points, superdomain_operator.num_extended_dofs);
        superdomain_operator.Pt.multiply(Tu_sup, tree.coarse_dofs);
    }
    timer.stop("subdomain.tree_construction.superdomain");
}
