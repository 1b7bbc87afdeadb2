/*
 * harness_main.c -- command-line harness around the drop-in attention() entry point.
 *
 * Written fresh for this engine; it keeps the reference harness' external contract
 * (attention-mpi.c:497-541): `prog <data file>`, data file = int32 m,n,dk,dv then fp64
 * Q, K, V and the expected m x dv answers (mpi.c:425-454,472-481), the elapsed time
 * covers exactly the attention() call (mpi.c:519-522), acceptance is |result-expected|
 * <= 0.02 for every element (mpi.c:476,483), stdout is "Correct!\nElapsed time: %.2lf us"
 * or "Wrong!" (mpi.c:526-532).  Differences, all on purpose:
 *   - 64-bit element counts and file offsets (the reference's `int offset`, mpi.c:472,
 *     overflows beyond 2 GiB of input, which the 8-GPU configuration exceeds);
 *   - the NaN test looks at the element being compared (mpi.c:483 tests column 1);
 *   - no MPI: rank/size come from RANK / WORLD_SIZE (torchrun-style launchers) and default
 *     to 0 / 1, where one process drives SDPA_NGPUS GPUs;
 *   - sdpa_runtime_init() plays the part of MPI_Init (outside the timed region);
 *   - the matrices are read into pinned memory (sdpa_host_alloc), so the transfers inside the timed
 *     call run at the PCIe rate; HARNESS_PAGEABLE=1 uses malloc like the reference (the library then
 *     stages the arrays through its pinned ring).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "../../include/sdpa_b200.h"

static void fail(const char* msg)
{
    fprintf(stderr, "%s\n", msg);
    exit(1);
}

static int g_pageable = 0;

static double* host_block(size_t count)
{
    const size_t bytes = sizeof(double) * (count ? count : 1);
    return (double*)(g_pageable ? malloc(bytes) : sdpa_host_alloc(bytes));
}

static void host_release(double* p)
{
    if (g_pageable) free(p);
    else sdpa_host_free(p);
}

static double* read_block(FILE* f, size_t count)
{
    double* p = host_block(count);
    if (!p) fail("Out of host memory.");
    if (fread(p, sizeof(double), count, f) != count) fail("Invalid testing data.");
    return p;
}

static int check_answers(const char* path, const double* result, int m, int n, int dk, int dv)
{
    FILE* f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "Cannot open answer file: %s\n", path);
        return 0;
    }
    const int64_t offset = 16 + 8 * ((int64_t)m * dk + (int64_t)n * dk + (int64_t)n * dv);
    if (fseeko(f, (off_t)offset, SEEK_SET) != 0) fail("Invalid testing data.");
    double* row = (double*)malloc(sizeof(double) * (size_t)(dv > 0 ? dv : 1));
    int ok = 1;
    for (int i = 0; i < m && ok; ++i) {
        if (fread(row, sizeof(double), (size_t)dv, f) != (size_t)dv) fail("Invalid testing data.");
        for (int j = 0; j < dv; ++j) {
            const double got = result[(size_t)i * dv + j];
            if (isnan(got) || fabs(got - row[j]) > 0.02) {
                printf("Expect result[%d][%d] to be %lf, but it is %lf\n", i, j, row[j], got);
                ok = 0;
                break;
            }
        }
    }
    free(row);
    fclose(f);
    return ok;
}

int main(int argc, char** argv)
{
    if (argc < 2) {
        fprintf(stderr, "Usage: %s <testing data>\n", argv[0]);
        return 1;
    }
    const char* er = getenv("RANK");
    const char* ew = getenv("WORLD_SIZE");
    const int rank = er ? atoi(er) : 0;
    const int size = ew ? atoi(ew) : 1;
    const char* ep = getenv("HARNESS_PAGEABLE");
    g_pageable = ep && *ep == '1';

    double *Q = NULL, *K = NULL, *V = NULL, *result = NULL;
    int m = 0, n = 0, dk = 0, dv = 0;
    if (rank == 0) {
        FILE* f = fopen(argv[1], "rb");
        if (!f) {
            fprintf(stderr, "Cannot open file: %s\n", argv[1]);
            return 1;
        }
        int32_t hdr[4];
        if (fread(hdr, sizeof(int32_t), 4, f) != 4) fail("Invalid testing data.");
        m = hdr[0]; n = hdr[1]; dk = hdr[2]; dv = hdr[3];
        if (m < 0 || n < 0 || dk < 1 || dv < 1) fail("Invalid testing data.");
        Q = read_block(f, (size_t)m * dk);
        K = read_block(f, (size_t)n * dk);
        V = read_block(f, (size_t)n * dv);
        fclose(f);
        result = host_block((size_t)m * dv);
        if (!result) fail("Out of host memory.");
    }

    if (sdpa_runtime_init(rank, size) != SDPA_OK) {
        fprintf(stderr, "sdpa_b200: %s\n", sdpa_last_error());
        return 1;
    }

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    attention(Q, K, V, result, m, n, dk, dv, rank, size);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double us = (double)(t1.tv_sec - t0.tv_sec) * 1e6 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-3;
    if (size > 1 && sdpa_runtime_max(&us) != SDPA_OK) {   /* the slowest rank's time, as MPI_Reduce(MAX) at mpi.c:524 */
        fprintf(stderr, "sdpa_b200: %s\n", sdpa_last_error());
        return 1;
    }

    int rc = 0;
    if (rank == 0) {
        if (check_answers(argv[1], result, m, n, dk, dv)) printf("Correct!\nElapsed time: %.2lf us\n", us);
        else { puts("Wrong!"); }
    }
    sdpa_runtime_shutdown();
    host_release(Q); host_release(K); host_release(V); host_release(result);
    return rc;
}
