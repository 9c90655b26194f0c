// kmeans.cu -- deterministic M-step sums for the global k-means of get_basis.py:210 (sklearn _k_means_lloyd.pyx:209-218).
//
// Each persistent CTA owns a contiguous range of points, streams it through shared memory with TMA bulk copies and adds
// the points, in ascending point order, into a private (K x D) fp32 table in shared memory: thread d owns column d, so
// there are no floating-point atomics and no bank conflicts.  The
// per-CTA tables are then reduced in CTA order in fp64.  The result does not depend on scheduling, and across GPU
// counts it differs only by the (fp64) order of the partial sums.
#include "common.cuh"

namespace gfs {

constexpr int KM_CH = 32;          // points per TMA chunk
constexpr int KM_THREADS = 288;    // warps 0-7: one thread per feature column; warp 8: TMA issue, label staging, counts

// X is row-major, so a chunk of 32 points is ONE contiguous cp.async.bulk (TMA) of 32*D*4 bytes; two chunks are in flight
// while the column threads add the current one into the shared-memory table.
__global__ void __launch_bounds__(KM_THREADS, 1)
kmeans_partial_kernel(const float* __restrict__ X, int64_t n, int D, const int32_t* __restrict__ labels, int K,
                      float* __restrict__ partial, int32_t* __restrict__ pcount) {
    extern __shared__ __align__(128) unsigned char sm_raw[];
    float* stg = reinterpret_cast<float*>(sm_raw);                          // [2][KM_CH][D]   (first: keeps 128 B alignment)
    float* tab = stg + 2 * KM_CH * D;                                       // [K][D]
    int* cnt = reinterpret_cast<int*>(tab + (size_t)K * D);                 // [K]
    int* lab = cnt + ((K + 3) & ~3);                                        // [2][KM_CH]
    uint64_t* full = reinterpret_cast<uint64_t*>(lab + 2 * KM_CH);          // [2]
    const int tid = threadIdx.x, lane = tid & 31;
    const bool helper = tid >= 256;

    for (int i = tid; i < K * D; i += KM_THREADS) tab[i] = 0.0f;
    for (int i = tid; i < K; i += KM_THREADS) cnt[i] = 0;
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    int64_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + KM_CH - 1) / KM_CH * KM_CH;
    const int64_t i0 = per * blockIdx.x;
    const int64_t i1 = (i0 + per) < n ? (i0 + per) : n;
    const int nchunks = i0 < i1 ? (int)((i1 - i0 + KM_CH - 1) / KM_CH) : 0;

    auto issue = [&](int c) {   // helper warp only
        if (c < nchunks) {
            const int64_t p0 = i0 + (int64_t)c * KM_CH;
            const int cc = (int)((i1 - p0) < KM_CH ? (i1 - p0) : KM_CH);
            lab[(c & 1) * KM_CH + lane] = lane < cc ? labels[p0 + lane] : 0;
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)cc * (uint32_t)D * 4u;
                mbar_arrive_expect_tx(&full[c & 1], bytes);
                tma_load_1d(stg + (size_t)(c & 1) * KM_CH * D, X + p0 * D, bytes, &full[c & 1]);
            }
        }
    };
    if (helper) {
        issue(0);
        issue(1);
    }
    __syncthreads();

    for (int c = 0; c < nchunks; ++c) {
        const int64_t p0 = i0 + (int64_t)c * KM_CH;
        const int cc = (int)((i1 - p0) < KM_CH ? (i1 - p0) : KM_CH);
        const int* L = lab + (c & 1) * KM_CH;
        mbar_wait(&full[c & 1], (c >> 1) & 1);
        if (!helper) {
            const int d = tid;
            if (d < D) {
                const float* S = stg + (size_t)(c & 1) * KM_CH * D + d;
                int j = 0;
                for (; j + 1 < cc; j += 2) {      // ascending point order; two independent read-modify-writes when labels differ
                    const int l0 = L[j], l1 = L[j + 1];
                    const float v0 = S[j * D], v1 = S[(j + 1) * D];
                    if (l0 != l1) {
                        const float a = tab[l0 * D + d], b = tab[l1 * D + d];
                        tab[l0 * D + d] = a + v0;
                        tab[l1 * D + d] = b + v1;
                    } else {
                        tab[l0 * D + d] = (tab[l0 * D + d] + v0) + v1;
                    }
                }
                if (j < cc) tab[L[j] * D + d] += S[j * D];
            }
        } else if (lane < cc) {
            atomicAdd(&cnt[L[lane]], 1);
        }
        __syncthreads();                          // chunk buffer and its labels are free again
        if (helper) issue(c + 2);
    }
    __syncthreads();
    float* o = partial + (size_t)blockIdx.x * K * D;
    for (int i = tid; i < K * D; i += KM_THREADS) o[i] = tab[i];
    for (int i = tid; i < K; i += KM_THREADS) pcount[(size_t)blockIdx.x * K + i] = cnt[i];
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
    GFS_REQUIRE(D <= 256 && D % 4 == 0, GFS_ERR_UNSUPPORTED, "gfs_kmeans_accumulate: D=%d must be a multiple of 4, <= 256", D);
    GFS_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0, GFS_ERR_BAD_ARG, "gfs_kmeans_accumulate: X must be 16-byte aligned");
    const size_t smem = (size_t)(2 * KM_CH + K) * D * sizeof(float) + (size_t)((K + 3) & ~3) * sizeof(int) + 2 * KM_CH * sizeof(int) + 64;
    GFS_REQUIRE(smem <= 220 * 1024, GFS_ERR_UNSUPPORTED, "gfs_kmeans_accumulate: K*D=%d does not fit shared memory", K * D);
    const int P = sm_count();
    GFS_REQUIRE(P > 0, GFS_ERR_CUDA, "gfs_kmeans_accumulate: cannot query the device");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(kmeans_partial_kernel), 220 * 1024));
    kmeans_partial_kernel<<<P, KM_THREADS, smem, st>>>(X, n, D, labels, K, partial, pcount);
    GFS_LAUNCH_OK("kmeans_partial_kernel");
    kmeans_reduce_kernel<<<(K * D + 255) / 256, 256, 0, st>>>(partial, pcount, P, K * D, K, sums, counts);
    GFS_LAUNCH_OK("kmeans_reduce_kernel");
    return GFS_OK;
}

namespace gfs {

// ---------------------------------------------------------------------------------------------------------------
// The rest of a Lloyd iteration as two launches (it used to be ~25 elementwise / reduction launches of the host library:
// at the sharded size of BASELINE.json configs[3], 500 k points per GPU, those fixed costs were half of the iteration).
//
//   kmeans_pack_kernel    packed = [sums (K*D, written by the reduce kernel) | counts as fp64 (K) | #labels changed (1)]
//                         -- the ONE buffer the sharded version all-reduces (get_basis.py:210)
//   kmeans_update_kernel  sklearn _k_means_common.pyx:_average_centers + the convergence quantities of _kmeans.py:_kmeans_single_lloyd:
//                         new = count > 0 ? sums / count : 0 (an empty cluster keeps the zero sum; the rare relocation is the
//                         caller's), shift = sum (new - old)^2 (fp64, fixed order), #empty; also the transposed (D, Kp) copy of
//                         the new centres that the next E-step reads.  result = [changed, shift, empty] (fp64) = one D2H read.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
kmeans_pack_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ labels_old, int64_t n, const int64_t* __restrict__ counts,
                   int K, int KD, double* __restrict__ packed, unsigned long long* __restrict__ scratch) {
    // scratch[0]: mismatch counter, scratch[1]: ticket; both are left at zero for the next call
    unsigned long long mine = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mine += labels[i] != labels_old[i];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&scratch[0], mine);          // integer sum: order does not matter
    __threadfence();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(&scratch[1], 1ull) == (unsigned long long)gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int c = threadIdx.x; c < K; c += blockDim.x) packed[KD + c] = (double)counts[c];
    if (threadIdx.x == 0) {
        packed[KD + K] = (double)atomicAdd(&scratch[0], 0ull);
        scratch[0] = 0ull;
        scratch[1] = 0ull;
    }
}

__global__ void __launch_bounds__(1024)
kmeans_update_kernel(const double* __restrict__ packed, const float* __restrict__ centers_old, int K, int D, int Kp,
                     float* __restrict__ centers_new, float* __restrict__ ct, double* __restrict__ result) {
    __shared__ double red[32];
    const int KD = K * D;
    double shift = 0.0;
    for (int i = threadIdx.x; i < KD; i += 1024) {
        const int c = i / D, d = i - c * D;
        const double cnt = rint(packed[KD + c]);
        const float nw = cnt > 0.0 ? (float)(packed[i] / cnt) : 0.0f;
        const double df = (double)(nw - centers_old[i]);
        shift += df * df;
        centers_new[i] = nw;
        ct[(int64_t)d * Kp + c] = nw;
    }
    double empty = 0.0;
    for (int c = threadIdx.x; c < K; c += 1024) empty += rint(packed[KD + c]) > 0.0 ? 0.0 : 1.0;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        shift += __shfl_xor_sync(0xffffffffu, shift, o);
        empty += __shfl_xor_sync(0xffffffffu, empty, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = shift;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += red[w];       // fixed order: deterministic
        result[1] = t;
        result[0] = packed[KD + K];
    }
    __syncthreads();
    if (lane == 0) red[warp] = empty;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += red[w];
        result[2] = t;
    }
}

}  // namespace gfs

extern "C" int gfs_kmeans_pack(const int32_t* labels, const int32_t* labels_old, int64_t n, const int64_t* counts, int K, int D,
                               double* packed, void* scratch16, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(labels && labels_old && counts && packed && scratch16 && n > 0 && K > 0 && D > 0, GFS_ERR_BAD_ARG, "gfs_kmeans_pack: bad argument");
    const int64_t want = (n + 256 * 8 - 1) / (256 * 8);
    const unsigned grid = (unsigned)(want < 1 ? 1 : want > 592 ? 592 : want);
    kmeans_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(labels, labels_old, n, counts, K, K * D, packed,
                                                                           static_cast<unsigned long long*>(scratch16));
    GFS_LAUNCH_OK("kmeans_pack_kernel");
    return GFS_OK;
}

extern "C" int gfs_kmeans_update(const double* packed, const float* centers_old, int K, int D, int Kp, float* centers_new, float* centers_t,
                                 double* result3, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(packed && centers_old && centers_new && centers_t && result3 && K > 0 && D > 0 && Kp >= K, GFS_ERR_BAD_ARG,
                "gfs_kmeans_update: bad argument");
    kmeans_update_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(packed, centers_old, K, D, Kp, centers_new, centers_t, result3);
    GFS_LAUNCH_OK("kmeans_update_kernel");
    return GFS_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// k-means++ seeding step (sklearn _kmeans.py:_kmeans_plusplus, as get_basis.py:210 reaches it through KMeans(init=
// 'k-means++')): for T <= 8 candidate centres at once,
//     m[t][i] = min( max(float(|x_i|^2 - 2 x_i.c_t + |c_t|^2), 0), closest[i] ),      pot[t] = sum_i m[t][i]   (fp64)
// sklearn evaluates these distances on float64 upcasts of the float32 data and rounds the result to float32
// (metrics/pairwise.py:_euclidean_distances_upcast); the kernel does the same -- fp64 fma chains, one rounding to fp32 --
// so the values agree with sklearn's bit for bit except where an fp64 summation-order difference straddles an fp32
// rounding boundary (~1e-9 of the values).  One thread per point over the channel-major copy the E-step already keeps
// resident: every channel is ONE coalesced load feeding T DFMAs against the candidates in shared memory, a single pass
// over X (HBM bound: n*D*4 bytes; 8 DFMA per 4 bytes stay below the B200's fp64 rate).
// ---------------------------------------------------------------------------------------------------------------
namespace gfs {

constexpr int PP_T = 8;
constexpr int PP_THREADS = 256;

__global__ void __launch_bounds__(PP_THREADS)
kmeans_pp_trial_kernel(const float* __restrict__ xt, int64_t npad, int64_t n, int D, const double* __restrict__ xsq,
                       const float* __restrict__ cand, int T, const float* __restrict__ closest, float* __restrict__ m_out,
                       double* __restrict__ pots) {
    extern __shared__ __align__(16) double csd[];   // [D][8] candidate coordinates, then [8] squared norms
    double* csq = csd + D * PP_T;
    for (int i = threadIdx.x; i < D * PP_T; i += PP_THREADS) {
        const int c = i / PP_T, t = i - c * PP_T;
        csd[i] = t < T ? (double)cand[(int64_t)t * D + c] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x < PP_T) {
        double s = 0.0;
        for (int c = 0; c < D; ++c) s = fma(csd[c * PP_T + threadIdx.x], csd[c * PP_T + threadIdx.x], s);
        csq[threadIdx.x] = s;
    }
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * PP_THREADS + threadIdx.x;
    const bool on = i < n;
    const int64_t ii = on ? i : 0;
    double dot[PP_T];
#pragma unroll
    for (int t = 0; t < PP_T; ++t) dot[t] = 0.0;
    const float* xp = xt + ii;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
        const double x = (double)__ldg(xp + (int64_t)c * npad);
        const double2 w0 = *reinterpret_cast<const double2*>(csd + c * PP_T), w1 = *reinterpret_cast<const double2*>(csd + c * PP_T + 2);
        const double2 w2 = *reinterpret_cast<const double2*>(csd + c * PP_T + 4), w3 = *reinterpret_cast<const double2*>(csd + c * PP_T + 6);
        dot[0] = fma(x, w0.x, dot[0]);
        dot[1] = fma(x, w0.y, dot[1]);
        dot[2] = fma(x, w1.x, dot[2]);
        dot[3] = fma(x, w1.y, dot[3]);
        dot[4] = fma(x, w2.x, dot[4]);
        dot[5] = fma(x, w2.y, dot[5]);
        dot[6] = fma(x, w3.x, dot[6]);
        dot[7] = fma(x, w3.y, dot[7]);
    }
    const double xs = xsq[ii];
    const float cl = closest ? closest[ii] : INFINITY;
    __shared__ double red[PP_T][PP_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < PP_T; ++t) {
        if (t < T) {                                  // warp-uniform
            const float d = fmaxf((float)((xs - 2.0 * dot[t]) + csq[t]), 0.0f);
            const float m = fminf(d, cl);
            if (on) m_out[(int64_t)t * npad + i] = m;
            double s = on ? (double)m : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[t][warp] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < T) {
        double s = 0.0;
        for (int w = 0; w < PP_THREADS / 32; ++w) s += red[threadIdx.x][w];
        atomicAdd(pots + threadIdx.x, s);
    }
}

}  // namespace gfs

extern "C" int gfs_kmeans_pp_trial(const float* xt, int64_t npad, int64_t n, int D, const double* xsq, const float* cand, int T,
                                   const float* closest, float* m_out, double* pots, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(xt && xsq && cand && m_out && pots, GFS_ERR_BAD_ARG, "gfs_kmeans_pp_trial: null pointer");
    GFS_REQUIRE(n > 0 && npad >= n && D > 0 && T > 0, GFS_ERR_BAD_ARG, "gfs_kmeans_pp_trial: non-positive size");
    GFS_REQUIRE(T <= PP_T, GFS_ERR_UNSUPPORTED, "gfs_kmeans_pp_trial: T=%d candidates > %d is not built", T, PP_T);
    GFS_REQUIRE(D <= 512, GFS_ERR_UNSUPPORTED, "gfs_kmeans_pp_trial: D=%d > 512 is not built", D);
    const size_t smem = (size_t)(D * PP_T + PP_T) * sizeof(double);
    const int64_t grid = (n + PP_THREADS - 1) / PP_THREADS;
    GFS_REQUIRE(grid < ((int64_t)1 << 31), GFS_ERR_UNSUPPORTED, "gfs_kmeans_pp_trial: n too large");
    kmeans_pp_trial_kernel<<<(unsigned)grid, PP_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(xt, npad, n, D, xsq, cand, T, closest,
                                                                                                m_out, pots);
    GFS_LAUNCH_OK("kmeans_pp_trial_kernel");
    return GFS_OK;
}
