// poisson.cpp -- the driver with the reference's command line and log lines
//   poisson <directory> <polynomial degree> <polynomial reduction> <subdomain overlap> <superdomain overlap> [solver_id]
// (/root/reference/poisson.cpp:40-81, 150-251).  Single process = single GPU; multi-GPU runs are launched one process
// per GPU with PRFDD_RANK / PRFDD_NRANKS / PRFDD_NCCL_ID_FILE in the environment (rank 0 writes the ncclUniqueId file).
#include <string>
#include "../../../include/prfdd_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <vector>

// The "Timings:" table of the reference's simulation_data() (poisson.cpp:253-401): same rows, same grouping of the
// Timer keys.  Printed when PRFDD_TIMINGS is set; phase timing fences the stream around every phase (as the
// reference's Timer does with device.finish()), so it is off by default and the solve then runs un-fenced.
static void timings_report(prfdd_solver *s)
{
    auto t = [&](const char *key) { const double v = prfdd_solver_timer_total(s, key); return v < 0.0 ? 0.0 : v; };
    auto sum = [&](std::initializer_list<const char *> keys) { double v = 0.0; for (auto k : keys) v += t(k); return v; };
    const double inner_products = t("domain.inner_products"), residual_norm = t("domain.residual_norm");
    const double vector_operations = t("domain.vector_operations"), operator_application = t("domain.operator_application");
    const double stitching = t("subdomain.stitching");
    const double subdomain_solver = sum({"subdomain.inner_products", "subdomain.residual_norm", "subdomain.vector_operations",
        "subdomain.operator_application", "subdomain.preconditioner.assemble_subdomain", "subdomain.preconditioner.assemble_composite",
        "subdomain.preconditioner.memcpy", "subdomain.preconditioner.vector_operations", "subdomain.preconditioner.down_leg_gpu",
        "subdomain.preconditioner.coarse_grid_solver", "subdomain.preconditioner.up_leg_gpu",
        "subdomain.preconditioner.unassemble_subdomain", "subdomain.preconditioner.unassemble_composite"});
    const double tree_construction = sum({"subdomain.tree_construction.gpu_to_gpu", "subdomain.tree_construction.subdomain",
        "subdomain.tree_construction.gpu_to_cpu", "subdomain.tree_construction.assemble_coarse", "subdomain.tree_construction.superdomain"});
    const double tree_exchange = sum({"subdomain.tree_exchange.superdomain", "subdomain.tree_exchange.subdomain", "subdomain.tree_exchange.cpu_to_gpu"});
    const double total = inner_products + residual_norm + vector_operations + operator_application + tree_construction + tree_exchange +
                         stitching + subdomain_solver;
    const double d = total > 0.0 ? total : 1.0;
    printf("\nTimings:\n-------------------------------------------------------------------------\n");
    printf("Total                 = %12.08f s ( %6.02f )\n", total, 100.0);
    printf("Inner products        = %12.08f s ( %6.02f )\n", inner_products, 100.0 * inner_products / d);
    printf("Residual norm         = %12.08f s ( %6.02f )\n", residual_norm, 100.0 * residual_norm / d);
    printf("Vector operations     = %12.08f s ( %6.02f )\n", vector_operations, 100.0 * vector_operations / d);
    printf("Operator application  = %12.08f s ( %6.02f )\n", operator_application, 100.0 * operator_application / d);
    printf("Tree construction     = %12.08f s ( %6.02f )\n", tree_construction, 100.0 * tree_construction / d);
    printf("Tree exchange         = %12.08f s ( %6.02f )\n", tree_exchange, 100.0 * tree_exchange / d);
    printf("Subdomain stitching   = %12.08f s ( %6.02f )\n", stitching, 100.0 * stitching / d);
    printf("Subdomain solver      = %12.08f s ( %6.02f )\n", subdomain_solver, 100.0 * subdomain_solver / d);
}

int main(int argc, char *argv[])
{
    if (argc < 6)
    {
        printf("ERROR: Use as 'poisson <directory> <polynomial degree> <polynomial reduction> <subdomain overlap> <superdomain overlap>'\n");
        return EXIT_SUCCESS; // the reference's quit() exits with EXIT_SUCCESS (config.hpp:57-62)
    }
    printf("----------------------------------------------------------------------------------\n");
    printf("|  Full domain decomposition with polynomial reduction -- B200-native build      |\n");
    printf("----------------------------------------------------------------------------------\n\n");
    prfdd_options opt;
    prfdd_options_default(&opt);
    opt.poly_degree = atoi(argv[2]);
    opt.poly_reduction = atoi(argv[3]);
    opt.subdomain_overlap = atoi(argv[4]);
    opt.superdomain_overlap = atoi(argv[5]);
    opt.verbose = 1;
    const int solver_id = argc > 6 ? atoi(argv[6]) : 1; // poisson.cpp:224 ships solver_id = 1 (GMRES); 0 = FCG
    if (getenv("PRFDD_TOLERANCE")) opt.outer_tolerance = atof(getenv("PRFDD_TOLERANCE"));
    if (getenv("PRFDD_NO_PRECONDITIONER")) opt.use_preconditioner = 0;
    if (getenv("PRFDD_INNER_FCG")) opt.preconditioner_type = 0;
    unsigned char id[128];
    if (getenv("PRFDD_NRANKS") && atoi(getenv("PRFDD_NRANKS")) > 1)
    {
        opt.num_procs = atoi(getenv("PRFDD_NRANKS"));
        opt.proc_id = atoi(getenv("PRFDD_RANK"));
        FILE *f = fopen(getenv("PRFDD_NCCL_ID_FILE"), "rb");
        if (!f || fread(id, 1, 128, f) != 128) { printf("ERROR: cannot read the ncclUniqueId file\n"); return EXIT_FAILURE; }
        fclose(f);
        opt.nccl_unique_id = id;
        prfdd_set_device(opt.proc_id % prfdd_device_count());
    }
    else
    {
        printf("Running with:\n- Mode: 'CUDA'\n- Device id: 0\n\n"); // poisson.cpp:126-139
        prfdd_set_device(0);
    }
    prfdd_solver *s = nullptr;
    int rc = prfdd_solver_create(&s, argv[1], &opt, nullptr);
    if (rc) { printf("ERROR: %s\n", prfdd_error_string(rc)); return EXIT_FAILURE; }
    prfdd_solver_setup_problem(s, 4);
    const bool timings = getenv("PRFDD_TIMINGS") != nullptr;
    if (timings) prfdd_solver_timer_total(s, "__enable__");
    int iters = 0, hl = 0;
    std::vector<double> hist(2048);
    rc = prfdd_solver_solve(s, solver_id, &iters, hist.data(), (int)hist.size(), &hl);
    if (rc) { printf("ERROR: %s\n", prfdd_error_string(rc)); return EXIT_FAILURE; }
    if (getenv("PRFDD_OUTPUT")) // poisson.cpp:233-235 (VISUALIZATION): u_star, f, u -> <name>_<rank>.vtk
    {
        rc = prfdd_solver_output(s, getenv("PRFDD_OUTPUT"));
        if (rc) { printf("ERROR: %s\n", prfdd_error_string(rc)); return EXIT_FAILURE; }
        if (opt.use_preconditioner) // Subdomain::output (subdomain.tpp:4648-4791): the rank's region -> <name>_subdomain_<rank>.vtk
        {
            rc = prfdd_solver_output_subdomain(s, (std::string(getenv("PRFDD_OUTPUT")) + "_subdomain").c_str());
            if (rc) { printf("ERROR: %s\n", prfdd_error_string(rc)); return EXIT_FAILURE; }
        }
    }
    if (opt.proc_id == 0)
    {
        printf("\nRun info:\n-------------------------------------------------------------------------\n");
        printf("Number of dimensions: %lld\n", prfdd_solver_query(s, PRFDD_Q_DIM));
        printf("Total number of elements: %lld\n", prfdd_solver_query(s, PRFDD_Q_NUM_TOTAL_ELEMENTS));
        printf("Polynomial degree: %d\n", opt.poly_degree);
        printf("Function ID: %d\n", 4);
        printf("Subdomain overlap: %d\n", opt.subdomain_overlap);
        printf("Superdomain overlap: %d\n", opt.superdomain_overlap);
        printf("Solver data precision: %s\n", "double");
        printf("Solver tolerance: %g\n", opt.outer_tolerance);
        printf("Solver type: \"%s\"\n", (solver_id == 0) ? "FCG" : "GMRES");
        printf("Preconditioner data precision: %s\n", "double");
        printf("Preconditioner tolerance: %g\n", opt.inner_tolerance);
        printf("Preconditioner type: \"%s\"\n", (opt.preconditioner_type == 0) ? "FCG" : "GMRES");
        printf("Iterations: %d\n", iters);
        if (timings) timings_report(s);
    }
    prfdd_solver_destroy(s);
    return EXIT_SUCCESS;
}
