/*
 * ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of every device kernel on the reference's hot path, written as
 * the loops OCCA's Serial backend executes: one @outer/@inner iteration == one loop
 * trip, `@shared` arrays == block-local arrays, reductions == 128-wide tree partials
 * per block followed by a serial host sum (the host sums are in oracle/prfdd_oracle.py).
 *
 * Follows /root/reference:
 *   domain.okl:5-264        (stiffness_matrix_1/2, initialize_arrays, residual_norm,
 *                            projection_inner_products, solution_and_residual_update,
 *                            inner_product_flexible, residual_and_search_update, inner_product)
 *   subdomain.okl:4-366     (variable-degree stiffness_matrix_1/2, weighted reductions,
 *                            casts, restriction_1/2/3)
 *   csr_matrix.okl:5-48     (multiply, multiply_range, multiply_weight)
 *   math.okl:5-35           (set_to_value, invert_vector_elements, vector_vector_addition,
 *                            vector_scaling)
 *   subdomain.tpp:21-33, 47-61, 74-78 and AMG/kernels.cu:25-94 (Chebyshev smoother pieces,
 *                            host branches) ; AMG/csr_matrix.cpp:114-125 (host matvec).
 *
 * DIM, DType=double, EType=double and BLOCK_SIZE=128 are OCCA compile-time defines in the
 * reference (domain.tpp:337-340); here DIM is the run-time argument `dim`, and the
 * POLY_DEGREE macro (subdomain.tpp:3886-3890, a *floating-point* table) is the argument
 * `poly_degree`.
 *
 * Pinned against the real reference kernels: oracle/build_ref.py translates the .okl files
 * where they lie into oracle/_ref/libref_okl_{2,3}d.so and tests/test_oracle_pin.py compares
 * every function below with them bit for bit.
 */
#include <stddef.h>

/* liboracle_omp.so is this file compiled with -fopenmp -DORACLE_OMP: the @outer loops run on the host threads, which is what
 * OCCA's OpenMP backend does with them (config.hpp:34-36).  Every loop trip owns its outputs and the block reductions stay serial
 * inside a block, so the results are bit-identical to the serial build (tests/test_oracle_pin.py checks both). */
#ifdef ORACLE_OMP
#define OMP_STR(x) #x
#define OMP_PRAGMA(x) _Pragma(OMP_STR(x))
#define OMP_FOR_N(n) OMP_PRAGMA(omp parallel for schedule(static) if ((n) > 8192))   /* small loops (coarse AMG levels) stay serial */
#else
#define OMP_FOR_N(n)
#endif

#define BLOCK_SIZE 128
/* liboracle_f32.so is this file compiled with -DORACLE_DTYPE=float: the reference's `Float float` (AMG/config.hpp:4,
 * config.hpp:19-20 PTYPE Float), used for the AMG smoother / matvec loops of the FP32 V-cycle */
#ifndef ORACLE_DTYPE
#define ORACLE_DTYPE double
#endif
typedef ORACLE_DTYPE DType;
typedef double EType;

/* ------------------------------------------------------------------ domain.okl */

void o_stiffness_matrix_1(DType **GDu, const DType *u, const DType *D_hat, const DType **G, const int num_points, const int poly_degree, const int dim)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        int n_x = poly_degree + 1;
        int n_xy = n_x * n_x;
        int num_elem_points = (dim == 2) ? n_x * n_x : n_x * n_x * n_x;

        int e = idx / num_elem_points;
        int v = idx % num_elem_points;

        if (dim == 2)
        {
            int i = v % n_x;
            int j = v / n_x;

            DType Du_1 = 0.0;
            DType Du_2 = 0.0;

            for (int k = 0; k < n_x; k++)
            {
                Du_1 += D_hat[k + i * n_x] * u[e * num_elem_points + (k + j * n_x)];
                Du_2 += D_hat[k + j * n_x] * u[e * num_elem_points + (i + k * n_x)];
            }

            GDu[0][idx] = G[0][idx] * Du_1 + G[2][idx] * Du_2;
            GDu[1][idx] = G[2][idx] * Du_1 + G[1][idx] * Du_2;
        }
        else
        {
            int i = v % n_x;
            int j = (v / n_x) % n_x;
            int k = v / n_xy;

            DType Du_1 = 0.0;
            DType Du_2 = 0.0;
            DType Du_3 = 0.0;

            for (int p = 0; p < n_x; p++)
            {
                Du_1 += D_hat[p + i * n_x] * u[e * num_elem_points + (p + j * n_x + k * n_xy)];
                Du_2 += D_hat[p + j * n_x] * u[e * num_elem_points + (i + p * n_x + k * n_xy)];
                Du_3 += D_hat[p + k * n_x] * u[e * num_elem_points + (i + j * n_x + p * n_xy)];
            }

            GDu[0][idx] = G[0][idx] * Du_1 + G[3][idx] * Du_2 + G[4][idx] * Du_3;
            GDu[1][idx] = G[3][idx] * Du_1 + G[1][idx] * Du_2 + G[5][idx] * Du_3;
            GDu[2][idx] = G[4][idx] * Du_1 + G[5][idx] * Du_2 + G[2][idx] * Du_3;
        }
    }
}

void o_stiffness_matrix_2(DType *Au, const DType **GDu, const DType *D_hat, const int num_points, const int poly_degree, const int dim)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        int n_x = poly_degree + 1;
        int n_xy = n_x * n_x;
        int num_elem_points = (dim == 2) ? n_x * n_x : n_x * n_x * n_x;

        int e = idx / num_elem_points;
        int v = idx % num_elem_points;

        if (dim == 2)
        {
            int i = v % n_x;
            int j = v / n_x;

            DType Au_1 = 0.0;
            DType Au_2 = 0.0;

            for (int k = 0; k < n_x; k++)
            {
                Au_1 += D_hat[i + k * n_x] * GDu[0][e * num_elem_points + (k + j * n_x)];
                Au_2 += D_hat[j + k * n_x] * GDu[1][e * num_elem_points + (i + k * n_x)];
            }

            Au[idx] = Au_1 + Au_2;
        }
        else
        {
            int i = v % n_x;
            int j = (v / n_x) % n_x;
            int k = v / n_xy;

            DType Au_1 = 0.0;
            DType Au_2 = 0.0;
            DType Au_3 = 0.0;

            for (int p = 0; p < n_x; p++)
            {
                Au_1 += D_hat[i + p * n_x] * GDu[0][e * num_elem_points + (p + j * n_x + k * n_xy)];
                Au_2 += D_hat[j + p * n_x] * GDu[1][e * num_elem_points + (i + p * n_x + k * n_xy)];
                Au_3 += D_hat[k + p * n_x] * GDu[2][e * num_elem_points + (i + j * n_x + p * n_xy)];
            }

            Au[idx] = Au_1 + Au_2 + Au_3;
        }
    }
}

void o_initialize_arrays(DType *u_k, DType *r_k, const DType *f, const int num_points)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        u_k[idx] = 0.0;
        r_k[idx] = f[idx];
    }
}

/* shared tree reduction of one 128-wide block, exactly the OKL loop nest */
static DType block_tree(DType *s)
{
    for (int alive = ((BLOCK_SIZE + 1) / 2); 0 < alive; alive /= 2)
        for (int item = 0; item < BLOCK_SIZE; ++item)
            if (item < alive) s[item] += s[item + alive];
    return s[0];
}

void o_residual_norm(DType *block, const DType *r_k, const DType *QQt_r_k, const DType *dirichlet_mask, const int num_points, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType r_norm[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_points)
                r_norm[item] = r_k[idx] * QQt_r_k[idx] * dirichlet_mask[idx];
            else
                r_norm[item] = 0.0;
        }
        block[group] = block_tree(r_norm);
    }
}

void o_projection_inner_products(DType *block, const DType *z_k, const DType *r_k, const DType *p_k, const DType *q_k, const int num_points, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType gamma_sum[BLOCK_SIZE];
        DType theta_sum[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_points)
            {
                gamma_sum[item] = z_k[idx] * r_k[idx];
                theta_sum[item] = p_k[idx] * q_k[idx];
            }
            else
            {
                gamma_sum[item] = 0.0;
                theta_sum[item] = 0.0;
            }
        }
        block[group] = block_tree(gamma_sum);
        block[group + num_blocks] = block_tree(theta_sum);
    }
}

void o_solution_and_residual_update(DType *u_k, DType *r_kp1, const DType *r_k, const DType *p_k, const DType *q_k, DType alpha_k, const int num_points)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        u_k[idx] += alpha_k * p_k[idx];
        r_kp1[idx] = r_k[idx] - alpha_k * q_k[idx];
    }
}

void o_inner_product_flexible(DType *block, const DType *r_k, const DType *r_kp1, const DType *z_k, const int num_points, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType theta_sum[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_points)
                theta_sum[item] = (r_kp1[idx] - r_k[idx]) * z_k[idx];
            else
                theta_sum[item] = 0.0;
        }
        block[group] = block_tree(theta_sum);
    }
}

void o_residual_and_search_update(DType *p_k, DType *r_k, const DType *z_k, const DType *r_kp1, DType beta_k, const int num_points)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        p_k[idx] = z_k[idx] + beta_k * p_k[idx];
        r_k[idx] = r_kp1[idx];
    }
}

void o_inner_product_mask(DType *block, const DType *u_k, const DType *v_k, const DType *dirichlet_mask, const int num_points, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType sum[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_points)
                sum[item] = u_k[idx] * v_k[idx] * dirichlet_mask[idx];
            else
                sum[item] = 0.0;
        }
        block[group] = block_tree(sum);
    }
}

/* --------------------------------------------------------------- subdomain.okl */

void o_sub_stiffness_matrix_1(DType **GDu, const DType *u, const DType **D_hat_ptr, const int *offset, const int *vert, const int *level, const DType **G, const int num_points, const DType *poly_degree, const int dim)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        int o = offset[idx];
        int v = vert[idx];
        int l = level[idx];
        int n_x = poly_degree[l] + 1;
        int n_xy = n_x * n_x;
        const DType *D_hat = D_hat_ptr[l];

        if (dim == 2)
        {
            int i = v % n_x;
            int j = v / n_x;

            DType Du_1 = 0.0;
            DType Du_2 = 0.0;

            for (int k = 0; k < n_x; k++)
            {
                Du_1 += D_hat[k + i * n_x] * u[o + (k + j * n_x)];
                Du_2 += D_hat[k + j * n_x] * u[o + (i + k * n_x)];
            }

            GDu[0][idx] = G[0][idx] * Du_1 + G[2][idx] * Du_2;
            GDu[1][idx] = G[2][idx] * Du_1 + G[1][idx] * Du_2;
        }
        else
        {
            int i = v % n_x;
            int j = (v / n_x) % n_x;
            int k = v / n_xy;

            DType Du_1 = 0.0;
            DType Du_2 = 0.0;
            DType Du_3 = 0.0;

            for (int p = 0; p < n_x; p++)
            {
                Du_1 += D_hat[p + i * n_x] * u[o + (p + j * n_x + k * n_xy)];
                Du_2 += D_hat[p + j * n_x] * u[o + (i + p * n_x + k * n_xy)];
                Du_3 += D_hat[p + k * n_x] * u[o + (i + j * n_x + p * n_xy)];
            }

            GDu[0][idx] = G[0][idx] * Du_1 + G[3][idx] * Du_2 + G[4][idx] * Du_3;
            GDu[1][idx] = G[3][idx] * Du_1 + G[1][idx] * Du_2 + G[5][idx] * Du_3;
            GDu[2][idx] = G[4][idx] * Du_1 + G[5][idx] * Du_2 + G[2][idx] * Du_3;
        }
    }
}

void o_sub_stiffness_matrix_2(DType *Au, const DType **GDu, const DType **D_hat_ptr, const int *offset, const int *vert, const int *level, const int num_points, const DType *poly_degree, const int dim)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        int o = offset[idx];
        int v = vert[idx];
        int l = level[idx];
        int n_x = poly_degree[l] + 1;
        int n_xy = n_x * n_x;
        const DType *D_hat = D_hat_ptr[l];

        if (dim == 2)
        {
            int i = v % n_x;
            int j = v / n_x;

            DType Au_1 = 0.0;
            DType Au_2 = 0.0;

            for (int k = 0; k < n_x; k++)
            {
                Au_1 += D_hat[i + k * n_x] * GDu[0][o + (k + j * n_x)];
                Au_2 += D_hat[j + k * n_x] * GDu[1][o + (i + k * n_x)];
            }

            Au[idx] = Au_1 + Au_2;
        }
        else
        {
            int i = v % n_x;
            int j = (v / n_x) % n_x;
            int k = v / n_xy;

            DType Au_1 = 0.0;
            DType Au_2 = 0.0;
            DType Au_3 = 0.0;

            for (int p = 0; p < n_x; p++)
            {
                Au_1 += D_hat[i + p * n_x] * GDu[0][o + (p + j * n_x + k * n_xy)];
                Au_2 += D_hat[j + p * n_x] * GDu[1][o + (i + p * n_x + k * n_xy)];
                Au_3 += D_hat[k + p * n_x] * GDu[2][o + (i + j * n_x + p * n_xy)];
            }

            Au[idx] = Au_1 + Au_2 + Au_3;
        }
    }
}

void o_sub_inner_product(DType *block, const DType *u, const DType *v, const int num_values, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType uv[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_values)
                uv[item] = u[idx] * v[idx];
            else
                uv[item] = 0.0;
        }
        block[group] = block_tree(uv);
    }
}

void o_sub_weighted_inner_product(DType *block, const DType *u, const DType *v, const DType *w, const int num_values, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType uv[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_values)
                uv[item] = u[idx] * v[idx] * w[idx];
            else
                uv[item] = 0.0;
        }
        block[group] = block_tree(uv);
    }
}

void o_sub_projection_inner_products(DType *block, const DType *z_k, const DType *r_k, const DType *p_k, const DType *q_k, const DType *weight, const int num_values, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType gamma_sum[BLOCK_SIZE];
        DType theta_sum[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_values)
            {
                gamma_sum[item] = z_k[idx] * r_k[idx] * weight[idx];
                theta_sum[item] = p_k[idx] * q_k[idx] * weight[idx];
            }
            else
            {
                gamma_sum[item] = 0.0;
                theta_sum[item] = 0.0;
            }
        }
        block[group] = block_tree(gamma_sum);
        block[group + num_blocks] = block_tree(theta_sum);
    }
}

void o_sub_search_update_inner_product(DType *block, const DType *r_k, const DType *r_kp1, const DType *z_k, const DType *weight, const int num_points, const int num_blocks)
{
    OMP_FOR_N(num_blocks)
    for (int group = 0; group < num_blocks; ++group)
    {
        DType theta_sum[BLOCK_SIZE];
        for (int item = 0; item < BLOCK_SIZE; ++item)
        {
            int idx = group * BLOCK_SIZE + item;
            if (idx < num_points)
                theta_sum[item] = (r_kp1[idx] - r_k[idx]) * z_k[idx] * weight[idx];
            else
                theta_sum[item] = 0.0;
        }
        block[group] = block_tree(theta_sum);
    }
}

void o_copy_from_domain_data(DType *u, const EType *v, const int num_points)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++) u[idx] = (DType)(v[idx]);
}

void o_copy_to_domain_data(EType *u, const DType *v, const int num_points)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++) u[idx] = (EType)(v[idx]);
}

void o_restriction_1(DType *Ju, const DType *J_cf, const DType *u, const int num_points, const int n_f, const int n_c, const int dim)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        int num_elem_points_fine = (dim == 2) ? n_f * n_f : n_f * n_f * n_f;
        int num_elem_points_coarse = (dim == 2) ? n_f * n_c : n_f * n_f * n_c;

        int e = idx / num_elem_points_coarse;
        int v = idx % num_elem_points_coarse;

        DType Ju_ij = 0.0;

        if (dim == 2)
        {
            int i = v % n_f;
            int j = v / n_f;

            for (int k = 0; k < n_f; k++) Ju_ij += J_cf[j + k * n_c] * u[(i + k * n_f) + e * num_elem_points_fine];

            Ju[(i + j * n_f) + e * num_elem_points_coarse] = Ju_ij;
        }
        else
        {
            int i = v % n_c;
            int j = (v / n_c) % n_f;
            int k = v / (n_c * n_f);

            for (int l = 0; l < n_f; l++) Ju_ij += J_cf[i + l * n_c] * u[(l + j * n_f + k * (n_f * n_f)) + e * num_elem_points_fine];

            Ju[(i + j * n_c + k * (n_c * n_f)) + e * num_elem_points_coarse] = Ju_ij;
        }
    }
}

void o_restriction_2(DType *Ju, const DType *J_cf, const DType *u, const int num_points, const int n_f, const int n_c, const int dim)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        int num_elem_points_fine = (dim == 2) ? n_f * n_c : n_f * n_f * n_c;
        int num_elem_points_coarse = (dim == 2) ? n_c * n_c : n_f * n_c * n_c;

        int e = idx / num_elem_points_coarse;
        int v = idx % num_elem_points_coarse;

        DType Ju_ij = 0.0;

        if (dim == 2)
        {
            int i = v % n_c;
            int j = v / n_c;

            for (int k = 0; k < n_f; k++) Ju_ij += u[(j * n_f + k) + e * num_elem_points_fine] * J_cf[k * n_c + i];

            Ju[(i + j * n_c) + e * num_elem_points_coarse] = Ju_ij;
        }
        else
        {
            int i = v % n_c;
            int j = (v / n_c) % n_c;
            int k = v / (n_c * n_c);

            for (int l = 0; l < n_f; l++) Ju_ij += J_cf[j + l * n_c] * u[(i + l * n_c + k * (n_c * n_f)) + e * num_elem_points_fine];

            Ju[(i + j * n_c + k * (n_c * n_c)) + e * num_elem_points_coarse] = Ju_ij;
        }
    }
}

void o_restriction_3(DType *Ju, const DType *J_cf, const DType *u, const int num_points, const int n_f, const int n_c)
{
    OMP_FOR_N(num_points)
    for (int idx = 0; idx < num_points; idx++)
    {
        int num_elem_points_fine = n_f * n_c * n_c;
        int num_elem_points_coarse = n_c * n_c * n_c;

        int e = idx / num_elem_points_coarse;
        int v = idx % num_elem_points_coarse;

        DType Ju_ij = 0.0;

        int i = v % n_c;
        int j = (v / n_c) % n_c;
        int k = v / (n_c * n_c);

        for (int l = 0; l < n_f; l++) Ju_ij += J_cf[k + l * n_c] * u[(i + j * n_c + l * (n_c * n_c)) + e * num_elem_points_fine];

        Ju[(i + j * n_c + k * (n_c * n_c)) + e * num_elem_points_coarse] = Ju_ij;
    }
}

/* -------------------------------------------------------------- csr_matrix.okl */

void o_csr_multiply(DType *Au, const int *A_ptr, const int *A_col, const DType *A_val, const DType *u, int n)
{
    OMP_FOR_N(n)
    for (int i = 0; i < n; i++)
    {
        DType Au_i = 0.0;
        for (int j = A_ptr[i]; j < A_ptr[i + 1]; j++) Au_i += A_val[j] * u[A_col[j]];
        Au[i] = Au_i;
    }
}

void o_csr_multiply_range(DType *Au, const int *A_ptr, const int *A_col, const DType *A_val, const DType *u, int row_start, int row_end)
{
    OMP_FOR_N(row_end)
    for (int i = row_start; i <= row_end; i++)
    {
        DType Au_i = 0.0;
        for (int j = A_ptr[i]; j < A_ptr[i + 1]; j++) Au_i += A_val[j] * u[A_col[j]];
        Au[i] = Au_i;
    }
}

void o_csr_multiply_weight(DType *Au, const int *A_ptr, const int *A_col, const DType *A_val, const DType *u, const DType *weight, int n)
{
    OMP_FOR_N(n)
    for (int i = 0; i < n; i++)
    {
        DType Au_i = 0.0;
        for (int j = A_ptr[i]; j < A_ptr[i + 1]; j++) Au_i += A_val[j] * u[A_col[j]];
        Au[i] = Au_i * weight[i];
    }
}

/* -------------------------------------------------------------------- math.okl */

void o_set_to_value(DType *u, DType alpha, int n, int offset)
{
    OMP_FOR_N(n)
    for (int i = 0; i < n; i++) u[i + offset] = alpha;
}

void o_invert_vector_elements(DType *u, int n)
{
    OMP_FOR_N(n)
    for (int i = 0; i < n; i++) u[i] = 1.0 / u[i];
}

void o_vector_vector_addition(DType *uv, const DType alpha, const DType *u, const DType beta, const DType *v, const int n)
{
    OMP_FOR_N(n)
    for (int i = 0; i < n; i++) uv[i] = alpha * u[i] + beta * v[i];
}

void o_vector_scaling(DType *au, const DType alpha, const DType *u, const int n)
{
    OMP_FOR_N(n)
    for (int i = 0; i < n; i++) au[i] = alpha * u[i];
}

/* ---------------------- AMG: host branches of subdomain.tpp:19-83, AMG/csr_matrix.cpp:114-125 */

/* y = alpha*A*x + beta*y   (AMG/csr_matrix.cpp:114-125) */
void o_amg_matvec(DType *y, const int *ptr, const int *col, const DType *val, const DType *x, DType alpha, DType beta, int num_rows)
{
    OMP_FOR_N(num_rows)
    for (int row = 0; row < num_rows; row++)
    {
        DType Ax = 0.0;
        for (int idx = ptr[row]; idx < ptr[row + 1]; idx++) Ax += val[idx] * x[col[idx]];
        y[row] = alpha * Ax + beta * y[row];
    }
}

/* Sr = S*(f - A u); w = alpha*Sr   (subdomain.tpp:21-33) */
void o_scaled_residual(DType *Sr, DType *w, const int *ptr, const int *col, const DType *val, const DType *u, const DType *f, const DType *S, DType alpha, int num_rows)
{
    OMP_FOR_N(num_rows)
    for (int row = 0; row < num_rows; row++)
    {
        DType Ax = 0.0;
        for (int idx = ptr[row]; idx < ptr[row + 1]; idx++) Ax += val[idx] * u[col[idx]];
        Sr[row] = S[row] * (f[row] - Ax);
        w[row] = alpha * Sr[row];
    }
}

/* v = D*(A*(D*w)); w = alpha*r + v   (subdomain.tpp:47-61) */
void o_polynomial_evaluation(DType *w, DType *v, const int *ptr, const int *col, const DType *val, const DType *r, const DType *D_val, DType alpha, int num_rows)
{
    OMP_FOR_N(num_rows)
    for (int row = 0; row < num_rows; row++)
    {
        DType tmp = 0.0;
        for (int idx = ptr[row]; idx < ptr[row + 1]; idx++) tmp += val[idx] * D_val[col[idx]] * w[col[idx]];
        v[row] = D_val[row] * tmp;
    }
    OMP_FOR_N(num_rows)
    for (int row = 0; row < num_rows; row++) w[row] = alpha * r[row] + v[row];
}

/* u += D*w   (subdomain.tpp:74-78) */
void o_update_field(DType *u, const DType *w, const DType *D_val, int size)
{
    OMP_FOR_N(size)
    for (int idx = 0; idx < size; idx++) u[idx] += D_val[idx] * w[idx];
}

/* ---------------------- AMG: device branches, AMG/kernels.cu:11-94 as loops */

void o_vector_set_to_value(DType *data, const DType value, const int size)
{
    OMP_FOR_N(size)
    for (int idx = 0; idx < size; idx++) data[idx] = value;
}

void o_main_scaled_residual(DType *Sr, DType *w, const DType *f_m_Au, const DType *S, const DType alpha, const int size)
{
    OMP_FOR_N(size)
    for (int idx = 0; idx < size; idx++)
    {
        Sr[idx] = S[idx] * f_m_Au[idx];
        w[idx] = alpha * Sr[idx];
    }
}

void o_main_polynomial_evaluation(DType *w, DType *v, const DType *r, const DType *D_val, const DType alpha, const int size)
{
    OMP_FOR_N(size)
    for (int idx = 0; idx < size; idx++)
    {
        v[idx] *= D_val[idx];
        w[idx] = alpha * r[idx] + v[idx];
    }
}

void o_main_update_field(DType *u, const DType *w, const DType *D_val, const int size)
{
    OMP_FOR_N(size)
    for (int idx = 0; idx < size; idx++) u[idx] += D_val[idx] * w[idx];
}

void o_vector_multiplication(DType *uv, const DType *u, const DType *v, const int size)
{
    OMP_FOR_N(size)
    for (int idx = 0; idx < size; idx++) uv[idx] = u[idx] * v[idx];
}

/* ---------------------- RHS random draws: the reference calls glibc rand() with the default
 * seed 1 per rank (domain.tpp:549-550, 572-573); srand(1) reproduces a fresh process. */
#include <stdlib.h>
void o_rand_fill(DType *out, int n, unsigned int seed)
{
    srand(seed);
    for (int i = 0; i < n; i++) out[i] = (DType)(rand()) / (DType)(RAND_MAX);
}

/* serial left-to-right host sum of block partials (e.g. domain.tpp:926) */
DType o_serial_sum(const DType *a, int n)
{
    DType s = 0.0;
    for (int i = 0; i < n; i++) s += a[i];
    return s;
}
