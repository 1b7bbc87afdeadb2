/*
 * sdpa_oracle.c -- CPU restatement of the reference attention path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * (or as the timed CPU baseline), never as a compute path of the engine.
 *
 * Parity pin: the reference ships no golden vectors,
 * so this restatement is pinned against the reference itself, compiled
 * unmodified from /root/reference into oracle/_ref/ (see oracle/Makefile):
 *   - oracle_attention_f64       vs  attention.c       (bit-for-bit on the
 *     fixtures in tests/golden/, see tests/test_oracle.py)
 *   - oracle_sharded_attention_f32 vs attention-mpi.c   (<= 2e-6, the two
 *     differ only in fp32 summation order: AVX-512 lanes vs scalar)
 *
 * Each function cites the reference lines it restates
 * (ser.c = attention.c, mpi.c = attention-mpi.c).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* Shard map: balanced contiguous partition of n rows over `size`     */
/* ranks; the first n % size ranks own one extra row.  mpi.c:19-27    */
/* ------------------------------------------------------------------ */
int oracle_owner_count(int n, int size, int rank)
{
    int q = n / size;
    int r = n % size;
    return rank < r ? q + 1 : q;
}

int oracle_owner_disp(int n, int size, int rank)
{
    int q = n / size;
    int r = n % size;
    return rank * q + (rank < r ? rank : r);
}

/* ------------------------------------------------------------------ */
/* Casts.  mpi.c:31-64 (fp64 -> fp32, round to nearest even) and       */
/* mpi.c:68-101 (fp32 -> fp64, exact).                                 */
/* ------------------------------------------------------------------ */
void oracle_cvt_d2f(float* dst, const double* src, size_t count)
{
    for (size_t i = 0; i < count; ++i) dst[i] = (float)src[i];
}

void oracle_cvt_f2d(double* dst, const float* src, size_t count)
{
    for (size_t i = 0; i < count; ++i) dst[i] = (double)src[i];
}

/* ------------------------------------------------------------------ */
/* Serial fp64 definition of the answer.  ser.c:20-75:                 */
/*   scale 1/sqrt(dk) (ser.c:23), score = dot * scale with the dot     */
/*   accumulated left to right (ser.c:33-42), max-subtracted softmax   */
/*   (ser.c:47-59), result[i][d] = sum_j p_j V[j][d] left to right     */
/*   (ser.c:65-71).  Same operation order => bit-identical to the      */
/*   compiled reference.                                               */
/* `row_begin/row_end` restrict the Q rows evaluated (used for the     */
/* row-subset checks at the large configs); pass 0, m for everything.  */
/* ------------------------------------------------------------------ */
void oracle_attention_f64_rows(const double* Q, const double* K, const double* V,
                               double* result, int m, int n, int dk, int dv,
                               int row_begin, int row_end)
{
    (void)m;
    const double scale = 1.0 / sqrt((double)dk);
#if defined(_OPENMP)
#pragma omp parallel
#endif
    {
        double* p = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
#if defined(_OPENMP)
#pragma omp for schedule(dynamic, 4)
#endif
        for (int i = row_begin; i < row_end; ++i) {
            const double* q = Q + (size_t)i * dk;
            for (int j = 0; j < n; ++j) {
                const double* k = K + (size_t)j * dk;
                double acc = 0.0;
                for (int t = 0; t < dk; ++t) acc += q[t] * k[t];
                p[j] = acc * scale;
            }
            double mx = p[0];
            for (int j = 1; j < n; ++j)
                if (p[j] > mx) mx = p[j];
            double denom = 0.0;
            for (int j = 0; j < n; ++j) {
                p[j] = exp(p[j] - mx);
                denom += p[j];
            }
            for (int j = 0; j < n; ++j) p[j] /= denom;
            double* out = result + (size_t)i * dv;
            for (int d = 0; d < dv; ++d) {
                double acc = 0.0;
                for (int j = 0; j < n; ++j) acc += p[j] * V[(size_t)j * dv + d];
                out[d] = acc;
            }
        }
        free(p);
    }
}

void oracle_attention_f64(const double* Q, const double* K, const double* V,
                          double* result, int m, int n, int dk, int dv)
{
    oracle_attention_f64_rows(Q, K, V, result, m, n, dk, dv, 0, m);
}

/* ------------------------------------------------------------------ */
/* One Q row against one K/V shard, fp32, single pass with a running   */
/* max / running sum and rescale-on-every-key.  mpi.c:168-189 and the  */
/* update order of SURVEY Appendix A:                                  */
/*   old = rmax; rmax = max(s, rmax); corr = expf(old - rmax);         */
/*   rsum = rsum*corr + expf(s - rmax);                                */
/*   contrib = (j>0 ? contrib*corr : 0) + expf(s - rmax) * V_j         */
/* The dot product is a plain left-to-right fp32 sum (the reference    */
/* sums 4x16 AVX-512 lanes, mpi.c:103-121 -- same value up to fp32     */
/* reassociation).  Empty shard => contrib 0, lmax -inf, lsum 0.       */
/* ------------------------------------------------------------------ */
void oracle_online_softmax_row_f32(float* contrib, float* lmax, float* lsum,
                                   const float* q, const float* K_local,
                                   const float* V_local, int n_local, int dk,
                                   int dv, float scale)
{
    float rmax = -INFINITY;
    float rsum = 0.0f;
    for (int d = 0; d < dv; ++d) contrib[d] = 0.0f;
    for (int j = 0; j < n_local; ++j) {
        const float* k = K_local + (size_t)j * dk;
        float acc = 0.0f;
        for (int t = 0; t < dk; ++t) acc += q[t] * k[t];
        const float s = acc * scale;
        const float old = rmax;
        if (s > rmax) rmax = s;
        const float corr = expf(old - rmax);
        const float w = expf(s - rmax);
        rsum = rsum * corr + w;
        const float* v = V_local + (size_t)j * dv;
        if (j > 0)
            for (int d = 0; d < dv; ++d) contrib[d] *= corr;
        for (int d = 0; d < dv; ++d) contrib[d] += w * v[d];
    }
    *lmax = rmax;
    *lsum = rsum;
}

/* Batch of rows -> per-row partial state (contrib, lmax, lsum); this is the
 * inner loop of mpi.c:333-338 for one rank. */
void oracle_online_softmax_partials_f32(float* contrib, float* lmax, float* lsum,
                                        const float* Qf, const float* K_local,
                                        const float* V_local, int rows, int n_local,
                                        int dk, int dv)
{
    const float scale = 1.0f / sqrtf((float)dk); /* mpi.c:208 */
#if defined(_OPENMP)
#pragma omp parallel for schedule(dynamic, 8)
#endif
    for (int b = 0; b < rows; ++b)
        oracle_online_softmax_row_f32(contrib + (size_t)b * dv, lmax + b, lsum + b,
                                      Qf + (size_t)b * dk, K_local, V_local, n_local,
                                      dk, dv, scale);
}

/* ------------------------------------------------------------------ */
/* Cross-shard merge of the partial softmax states, for `rows` rows     */
/* and `shards` shards.  mpi.c:340-380 / SURVEY 3.3:                    */
/*   gmax = max_r lmax_r                      (allreduce MAX, :342)    */
/*   c_r  = expf(lmax_r - gmax)               (:347)                   */
/*   gsum = sum_r lsum_r * c_r                (allreduce SUM, :348,354)*/
/*   out  = sum_r contrib_r * c_r * inv,  inv = gsum==0 ? 0 : 1/gsum   */
/*          (each shard scales before the reduce, :358-362, :380)      */
/* Layout: contrib[shard][row][dv], lmax/lsum[shard][row].             */
/* The sum over shards runs in rank order (MPI leaves it unspecified). */
/* ------------------------------------------------------------------ */
void oracle_merge_partials_f32(float* out, const float* contrib, const float* lmax,
                               const float* lsum, int shards, int rows, int dv)
{
    for (int b = 0; b < rows; ++b) {
        float gmax = -INFINITY;
        for (int r = 0; r < shards; ++r) {
            float v = lmax[(size_t)r * rows + b];
            if (v > gmax) gmax = v;
        }
        float gsum = 0.0f;
        for (int r = 0; r < shards; ++r) {
            float c = expf(lmax[(size_t)r * rows + b] - gmax);
            gsum += lsum[(size_t)r * rows + b] * c;
        }
        const float inv = (gsum == 0.0f) ? 0.0f : 1.0f / gsum;
        float* o = out + (size_t)b * dv;
        for (int d = 0; d < dv; ++d) o[d] = 0.0f;
        for (int r = 0; r < shards; ++r) {
            const float c = expf(lmax[(size_t)r * rows + b] - gmax);
            const float* cr = contrib + ((size_t)r * rows + b) * dv;
            for (int d = 0; d < dv; ++d) o[d] += (cr[d] * c) * inv;
        }
    }
}

/* ------------------------------------------------------------------ */
/* Whole attention() of the MPI flavour emulated in one process for    */
/* `shards` ranks: d2f casts (mpi.c:224-225,303), owner partition      */
/* (mpi.c:199,236), per-shard online softmax (mpi.c:333-338), merge    */
/* (mpi.c:340-380), f2d write-back (mpi.c:373,396).  The Q batching    */
/* (B=512, mpi.c:200,307) does not change any value and is omitted.    */
/* Returns 0, or -1 on allocation failure.                             */
/* ------------------------------------------------------------------ */
int oracle_sharded_attention_f32(const double* Q, const double* K, const double* V,
                                 double* result, int m, int n, int dk, int dv,
                                 int shards)
{
    if (shards < 1) return -1;
    float* Qf = (float*)malloc(sizeof(float) * ((size_t)m * dk + 1));
    float* Kf = (float*)malloc(sizeof(float) * ((size_t)n * dk + 1));
    float* Vf = (float*)malloc(sizeof(float) * ((size_t)n * dv + 1));
    float* contrib = (float*)malloc(sizeof(float) * ((size_t)shards * m * dv + 1));
    float* lmax = (float*)malloc(sizeof(float) * ((size_t)shards * m + 1));
    float* lsum = (float*)malloc(sizeof(float) * ((size_t)shards * m + 1));
    float* outf = (float*)malloc(sizeof(float) * ((size_t)m * dv + 1));
    int rc = -1;
    if (Qf && Kf && Vf && contrib && lmax && lsum && outf) {
        oracle_cvt_d2f(Qf, Q, (size_t)m * dk);
        oracle_cvt_d2f(Kf, K, (size_t)n * dk);
        oracle_cvt_d2f(Vf, V, (size_t)n * dv);
        for (int r = 0; r < shards; ++r) {
            const int cnt = oracle_owner_count(n, shards, r);
            const int dsp = oracle_owner_disp(n, shards, r);
            oracle_online_softmax_partials_f32(
                contrib + (size_t)r * m * dv, lmax + (size_t)r * m, lsum + (size_t)r * m,
                Qf, Kf + (size_t)dsp * dk, Vf + (size_t)dsp * dv, m, cnt, dk, dv);
        }
        oracle_merge_partials_f32(outf, contrib, lmax, lsum, shards, m, dv);
        oracle_cvt_f2d(result, outf, (size_t)m * dv);
        rc = 0;
    }
    free(Qf); free(Kf); free(Vf); free(contrib); free(lmax); free(lsum); free(outf);
    return rc;
}

int oracle_num_threads(void)
{
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
