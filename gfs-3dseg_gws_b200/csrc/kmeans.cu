// kmeans.cu -- deterministic M-step sums for the global k-means of get_basis.py:210 (sklearn _k_means_lloyd.pyx:209-218).
//
// Each persistent CTA owns a contiguous range of points and adds them, in ascending point order, into a private
// (K x D) fp32 table in shared memory: thread d owns column d, so there are no atomics and no bank conflicts.  The
// per-CTA tables are then reduced in CTA order in fp64.  The result does not depend on scheduling, and across GPU
// counts it differs only by the (fp64) order of the partial sums.
#include "common.cuh"

namespace gfs {

__global__ void __launch_bounds__(256)
kmeans_partial_kernel(const float* __restrict__ X, int64_t n, int D, const int32_t* __restrict__ labels, int K,
                      float* __restrict__ partial, int32_t* __restrict__ pcount) {
    extern __shared__ float tab[];                       // [K][D]
    int* cnt = reinterpret_cast<int*>(tab + (size_t)K * D);   // [K]
    const int d = threadIdx.x;
    for (int i = d; i < K * D; i += 256) tab[i] = 0.0f;
    for (int i = d; i < K; i += 256) cnt[i] = 0;
    __syncthreads();
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t i0 = per * blockIdx.x;
    const int64_t i1 = (i0 + per) < n ? (i0 + per) : n;
    if (d < D) {
        int64_t i = i0;
        for (; i + 4 <= i1; i += 4) {
            int l[4];
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                l[u] = labels[i + u];
                v[u] = X[(i + u) * D + d];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) tab[l[u] * D + d] += v[u];
        }
        for (; i < i1; ++i) tab[labels[i] * D + d] += X[i * D + d];
    } else if (d == 255) {
        for (int64_t i = i0; i < i1; ++i) cnt[labels[i]] += 1;
    }
    __syncthreads();
    float* o = partial + (size_t)blockIdx.x * K * D;
    for (int i = d; i < K * D; i += 256) o[i] = tab[i];
    for (int i = d; i < K; i += 256) pcount[(size_t)blockIdx.x * K + i] = cnt[i];
}

__global__ void kmeans_reduce_kernel(const float* __restrict__ partial, const int32_t* __restrict__ pcount, int P, int KD, int K,
                                     double* __restrict__ sums, int64_t* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < KD) {
        double acc = 0.0;
        for (int p = 0; p < P; ++p) acc += (double)partial[(size_t)p * KD + i];
        sums[i] = acc;
    }
    if (i < K) {
        int64_t c = 0;
        for (int p = 0; p < P; ++p) c += pcount[(size_t)p * K + i];
        counts[i] = c;
    }
}

}  // namespace gfs

extern "C" int gfs_kmeans_partials(void) { return gfs::sm_count(); }

extern "C" int gfs_kmeans_accumulate(const float* X, int64_t n, int D, const int32_t* labels, int K, float* partial,
                                     int32_t* pcount, double* sums, int64_t* counts, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(X && labels && partial && pcount && sums && counts, GFS_ERR_BAD_ARG, "gfs_kmeans_accumulate: null pointer");
    GFS_REQUIRE(n > 0 && D > 0 && K > 0, GFS_ERR_BAD_ARG, "gfs_kmeans_accumulate: non-positive size");
    GFS_REQUIRE(D <= 254, GFS_ERR_UNSUPPORTED, "gfs_kmeans_accumulate: D=%d > 254 is not built", D);
    const size_t smem = (size_t)K * D * sizeof(float) + (size_t)K * sizeof(int);
    GFS_REQUIRE(smem <= 220 * 1024, GFS_ERR_UNSUPPORTED, "gfs_kmeans_accumulate: K*D=%d does not fit shared memory", K * D);
    const int P = sm_count();
    GFS_REQUIRE(P > 0, GFS_ERR_CUDA, "gfs_kmeans_accumulate: cannot query the device");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(kmeans_partial_kernel), 220 * 1024));
    kmeans_partial_kernel<<<P, 256, smem, st>>>(X, n, D, labels, K, partial, pcount);
    GFS_LAUNCH_OK("kmeans_partial_kernel");
    kmeans_reduce_kernel<<<(K * D + 255) / 256, 256, 0, st>>>(partial, pcount, P, K * D, K, sums, counts);
    GFS_LAUNCH_OK("kmeans_reduce_kernel");
    return GFS_OK;
}
