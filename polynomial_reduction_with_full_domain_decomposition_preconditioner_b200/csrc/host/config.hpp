// config.hpp -- types, macros and process-wide globals shared by all host classes.
// Mirrors /root/reference/config.hpp:19-65 (STYPE/PTYPE, BLOCK_SIZE, rstdout/pstdout, the globals
// dim / proc_id / num_procs / device / timer, quit()); the OCCA, MPI and HYPRE includes are gone.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <iostream>
#include <unordered_map>
#include "device.hpp"
#include "timer.hpp"
#include "comm.hpp"

typedef double Float;       // AMG/config.hpp:4
#define STYPE double        // config.hpp:19
#define PTYPE Float         // config.hpp:20
#ifndef BLOCK_SIZE
#define BLOCK_SIZE 128      // config.hpp:38-40
#endif

namespace prfdd_host
{
// the reference's globals (config.hpp:48-65), kept process-wide but inside a namespace so the
// shared library does not export symbols called `dim` or `timer`
extern int dim;
extern int proc_id;
extern int num_procs;
extern int verbose;
extern dev::device device;
extern Timer<double> timer;
extern Comm comm_world;      // stands where MPI_COMM_WORLD stands
extern FILE *pstdout_file;
} // namespace prfdd_host

// set-up stopwatch: with PRFDD_SETUP_TIMING set, rank 0 prints the seconds since the previous mark (host wall clock; set-up is host work)
inline void setup_mark(const char *what)
{
    static const bool on = getenv("PRFDD_SETUP_TIMING") != nullptr;
    static double last = -1.0;
    if (!on || prfdd_host::proc_id != 0) return;
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if (last >= 0.0 && what) fprintf(stderr, "[setup] %-44s %7.2f s\n", what, now - last);
    last = now;
}

// nested stopwatch for the stages inside one mark (own clock, so that the outer marks still measure whole stages)
inline void setup_submark(const char *what)
{
    static const bool on = getenv("PRFDD_SETUP_TIMING") != nullptr;
    static double last = -1.0;
    if (!on || prfdd_host::proc_id != 0) return;
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    const double now = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if (last >= 0.0 && what) fprintf(stderr, "[setup]     . %-40s %7.2f s\n", what, now - last);
    last = now;
}

#define rstdout(...) { if (prfdd_host::proc_id == 0 && prfdd_host::verbose) { printf(__VA_ARGS__); fflush(stdout); } }
#define pstdout(...) { if (prfdd_host::pstdout_file) { fprintf(prfdd_host::pstdout_file, __VA_ARGS__); fflush(prfdd_host::pstdout_file); } }
