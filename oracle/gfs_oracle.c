/*
 * oracle/gfs_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never measured
 * except as bench.py's cpu_baseline / --impl reference leg).
 *
 * Plain-C restatement of the integer/index-producing steps of the hot path with a
 * *pinned evaluation order*, so that the CUDA kernels can be checked bit-for-bit:
 *
 *   gfs_oracle_knn          <- model/dgcnn.py:17-23  (knn: -xx - inner - xx^T, topk)
 *   gfs_oracle_kmeans_assign<- sklearn 1.9.0 _k_means_lloyd.pyx:196-218 as called from
 *                              get_basis.py:210 (||c||^2 - 2 x.c, argmin, strict <)
 *   gfs_oracle_kmeans_accumulate <- sklearn _k_means_lloyd.pyx M-step (sum / count),
 *                              restated as an ascending-index sequential sum
 *
 * Pinned order (identical in csrc/knn.cu and csrc/kmeans.cu):
 *   dot(i,j)  = fma chain over channels c = 0..C-1 ascending, starting from +0.0f
 *   xx(i)     = dot(i,i)
 *   knn  d    = fmaf(2, dot, -xx_i) - xx_j      (== (-xx_i - (-2 dot)) - xx_j, one
 *               rounding each, because 2*dot is exact)
 *   knn order = larger d first, ties -> smaller j first
 *   kmeans s  = fmaf(-2, dot(x,c), cc)  with cc = dot(c,c); argmin, ties -> smaller c
 *
 * The reference's own matmul (MKL/cuBLAS) uses an unspecified summation order, so
 * reference-vs-oracle differences are confined to near-ties; tests/ classify them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float dot_chain(const float* a, long sa, const float* b, long sb, int C) {
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) acc = fmaf(a[c * sa], b[c * sb], acc);
    return acc;
}

/* x: (B, C, N) fp32 channel-major.  idx_out: (B, N, k) int32, dist_out (optional): (B, N, k). */
int gfs_oracle_knn(const float* x, int B, int C, int N, int k, int32_t* idx_out, float* dist_out) {
    if (k > N || k <= 0) return 1;
#pragma omp parallel
    {
        float* bd = (float*)malloc(sizeof(float) * (size_t)k);
        int32_t* bi = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
        float* xx = (float*)malloc(sizeof(float) * (size_t)N);
#pragma omp for schedule(static) collapse(1)
        for (int b = 0; b < B; ++b) {
            const float* xb = x + (size_t)b * C * N;
            for (int i = 0; i < N; ++i) xx[i] = dot_chain(xb + i, N, xb + i, N, C);
            for (int i = 0; i < N; ++i) {
                int n = 0;
                for (int j = 0; j < N; ++j) {
                    float dot = dot_chain(xb + i, N, xb + j, N, C);
                    float d = fmaf(2.0f, dot, -xx[i]) - xx[j];
                    /* insertion into (d desc, j asc); j ascending so ties never displace */
                    if (n == k && !(d > bd[k - 1])) continue;
                    int p = (n < k) ? n : k - 1;
                    while (p > 0 && d > bd[p - 1]) { bd[p] = bd[p - 1]; bi[p] = bi[p - 1]; --p; }
                    bd[p] = d; bi[p] = j;
                    if (n < k) ++n;
                }
                size_t o = ((size_t)b * N + i) * k;
                for (int t = 0; t < k; ++t) {
                    idx_out[o + t] = bi[t];
                    if (dist_out) dist_out[o + t] = bd[t];
                }
            }
        }
        free(bd); free(bi); free(xx);
    }
    return 0;
}

/* X: (n, D) row-major, Cc: (K, D) row-major.  labels: (n,) int32; best (optional): (n,) score. */
int gfs_oracle_kmeans_assign(const float* X, long n, int D, const float* Cc, int K,
                             int32_t* labels, float* best) {
    float* cc = (float*)malloc(sizeof(float) * (size_t)K);
    for (int c = 0; c < K; ++c) cc[c] = dot_chain(Cc + (size_t)c * D, 1, Cc + (size_t)c * D, 1, D);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const float* xi = X + (size_t)i * D;
        float bs = INFINITY; int bl = 0;
        for (int c = 0; c < K; ++c) {
            float dot = dot_chain(xi, 1, Cc + (size_t)c * D, 1, D);
            float s = fmaf(-2.0f, dot, cc[c]);
            if (s < bs) { bs = s; bl = c; }
        }
        labels[i] = bl;
        if (best) best[i] = bs;
    }
    free(cc);
    return 0;
}

/* sums: (K, D) fp64 exact-ish accumulation in ascending point order; counts: (K,) int64 */
int gfs_oracle_kmeans_accumulate(const float* X, long n, int D, const int32_t* labels, int K,
                                 double* sums, int64_t* counts) {
    memset(sums, 0, sizeof(double) * (size_t)K * D);
    memset(counts, 0, sizeof(int64_t) * (size_t)K);
    for (long i = 0; i < n; ++i) {
        int l = labels[i];
        if (l < 0 || l >= K) return 1;
        counts[l] += 1;
        for (int d = 0; d < D; ++d) sums[(size_t)l * D + d] += (double)X[(size_t)i * D + d];
    }
    return 0;
}
