// element.hpp -- per-element host record, field for field the reference's Element<DType>
// (/root/reference/element.hpp:18-55, element.tpp:5-41).
#pragma once
#include <cmath>
#include <set>
#include <vector>

#ifndef NUM_GEOM_FACTS
#define NUM_GEOM_FACTS 6
#endif

template <typename DType>
class Element
{
  public:
    // Descriptors
    int id;
    int dim;
    int poly_degree;
    int num_points;
    int offset;

    // Mesh
    int n_x, n_y, n_z;
    std::vector<DType> x, y, z;

    // Dirichlet boundary conditions
    std::vector<DType> dirichlet_mask;

    // Geometric factor
    std::vector<DType> geom_fact[NUM_GEOM_FACTS];

    // Connectivity
    std::vector<int> loc_num;
    std::vector<long long> glo_num;
    std::vector<long long> dof_num;
    std::vector<std::set<int>> vert_conn;
    std::vector<std::set<int>> edge_conn;
    std::vector<std::set<int>> face_conn;

    Element(int id_, int dim_, int poly_degree_) : id(id_), dim(dim_), poly_degree(poly_degree_), offset(0)
    {
        num_points = 1;
        for (int d = 0; d < dim; d++) num_points *= (poly_degree + 1);
        n_x = n_y = n_z = poly_degree + 1;
        x.resize(num_points);
        y.resize(num_points);
        z.resize(num_points);
        dirichlet_mask.resize(num_points);
        for (int g = 0; g < NUM_GEOM_FACTS; g++) geom_fact[g].resize(num_points);
        loc_num.resize(num_points);
        glo_num.resize(num_points);
        dof_num.resize(num_points);
        vert_conn.resize((dim == 2) ? 4 : 8);
        edge_conn.resize((dim == 2) ? 4 : 12);
        face_conn.resize((dim == 2) ? 0 : 6);
    }
    ~Element() {}
};
