// knn.cu -- fused fp32 kNN graph (replaces model/dgcnn.py:17-23: matmul + 4 elementwise passes + topk).
//
// One CTA owns 64 query points of one block and streams all N candidates in tiles of 128 through shared memory
// (cp.async, double buffered).  Distances are formed in registers by the shared FFMA core in the pinned order
//     dot = fma chain over c ascending;  d = fmaf(2, dot, -xx_i) - xx_j
// and dropped into a 64x128 shared tile; each warp then scans its 16 rows against the row's threshold (the k-th best at
// the last merge: a one-compare filter that rejects ~97 % of candidates after the first tile) and appends survivors to a
// 32-entry per-row buffer.  Only when the buffer would overflow (about 4 times per row at N=2048) is it merged into the
// row's sorted list: a warp-wide bitonic sort of 64-bit keys (orderable distance << 32 | ~index) followed by a bitonic
// top-32 merge -- 20 compare-exchange stages, no serial insertion.  Order: larger d first, ties -> smaller index.
// The N x N matrix never leaves the SM.
#include "fp32_tile.cuh"

namespace gfs {

constexpr int KNN_KC = 32;   // channels per pipeline stage

struct KnnSmem {
    float Bs[2][KNN_KC * T_COLS];   // 32 KB
    float Ds[T_ROWS * T_COLS];      // 32 KB  distance tile of the current candidate tile
    uint2 list[T_ROWS * 32];        // 16 KB  per query: 32 best so far, sorted descending; (x = orderable distance, y = ~index)
    uint2 buf[T_ROWS * 32];         // 16 KB  per query: raw (distance bits, index) appended since the last merge
    float tau[T_ROWS];              // distance of the k-th best at the last merge (filter threshold, never decreases)
    int fill[T_ROWS];
    // followed by As[C][64]
};

typedef unsigned long long u64;

// fp32 -> uint32 whose unsigned order equals the float order
__device__ __forceinline__ uint32_t ord_key(float d) {
    const uint32_t u = __float_as_uint(d);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_val(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// compare-exchange with the lane `lane ^ j2`: keep the larger key if keep_max else the smaller
__device__ __forceinline__ u64 cmpex(u64 k, int j2, bool keep_max) {
    const u64 o = __shfl_xor_sync(0xffffffffu, k, j2);
    return (keep_max == (k > o)) ? k : o;
}

// Merge the row's append buffer into its sorted list (one warp; key = (orderable d << 32) | ~index, larger = better, so
// equal distances order by ascending index).  Bitonic sort of the <=32 buffered keys, then the classic
// "max(list, reverse(sorted buffer))" + 5-stage bitonic merge keeps the 32 largest of the union, sorted descending.
__device__ __forceinline__ float knn_flush(KnnSmem& s, int q, int lane, int fill, int k) {
    const uint2 raw = s.buf[q * 32 + lane];
    u64 key = lane < fill ? (((u64)ord_key(__uint_as_float(raw.x)) << 32) | (u64)(~raw.y)) : 0ull;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
            const bool lower = (lane & j2) == 0;
            const bool desc = (lane & k2) == 0;
            key = cmpex(key, j2, lower == desc);
        }
    }
    const uint2 l2 = s.list[q * 32 + lane];
    const u64 lst = ((u64)l2.x << 32) | (u64)l2.y;
    const u64 rev = __shfl_sync(0xffffffffu, key, 31 - lane);
    key = lst > rev ? lst : rev;
#pragma unroll
    for (int j2 = 16; j2 > 0; j2 >>= 1) key = cmpex(key, j2, (lane & j2) == 0);
    s.list[q * 32 + lane] = make_uint2((uint32_t)(key >> 32), (uint32_t)key);
    const uint32_t kth = __shfl_sync(0xffffffffu, (uint32_t)(key >> 32), k - 1);
    const float tau = kth ? ord_val(kth) : -INFINITY;
    if (lane == 0) s.tau[q] = tau;
    return tau;
}

__global__ void __launch_bounds__(T_THREADS, 2)
knn_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, int k, const float* __restrict__ sqnorm,
           int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KnnSmem& s = *reinterpret_cast<KnnSmem*>(smem_raw);
    float* As = reinterpret_cast<float*>(smem_raw + sizeof(KnnSmem));

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.y, q0 = blockIdx.x * T_ROWS;
    const float* xb = x + (int64_t)b * bstride;
    const float* xxb = sqnorm + (int64_t)b * N;
    const unsigned lt_mask = (1u << lane) - 1u;

    const int nch = (C + KNN_KC - 1) / KNN_KC;
    const int ntiles = (N + T_COLS - 1) / T_COLS;
    const int S = ntiles * nch;

    for (int i = tid; i < T_ROWS * 32; i += T_THREADS) s.list[i] = make_uint2(0u, 0u);
    if (tid < T_ROWS) {
        s.tau[tid] = -INFINITY;
        s.fill[tid] = 0;
    }

    // query panel (all channels) + stage 0
    load_panel_async(As, T_ROWS, xb, N, C, q0, N, tid);
    {
        const int c1 = C < KNN_KC ? C : KNN_KC;
        load_panel_async(s.Bs[0], T_COLS, xb, N, c1, 0, N, tid);
    }
    cp_async_commit();

    float xq[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int q = q0 + ty * 8 + r;
        xq[r] = q < N ? xxb[q] : 0.0f;
    }

    float acc[8][8];
    for (int st = 0; st < S; ++st) {
        const int t = st / nch, ch = st - t * nch;
        if (st + 1 < S) {
            const int t1 = (st + 1) / nch, ch1 = (st + 1) - t1 * nch;
            const int c0 = ch1 * KNN_KC;
            const int cn = (C - c0) < KNN_KC ? (C - c0) : KNN_KC;
            load_panel_async(s.Bs[(st + 1) & 1], T_COLS, xb + (int64_t)c0 * N, N, cn, t1 * T_COLS, N, tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (ch == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;
        }
        const int c0 = ch * KNN_KC;
        const int cn = (C - c0) < KNN_KC ? (C - c0) : KNN_KC;
        tile_fma(As + c0 * T_ROWS, s.Bs[st & 1], cn, ty, tx, acc);

        if (ch == nch - 1) {
            const int j0 = t * T_COLS;
            float xc[8];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int j = j0 + h * 64 + tx * 4 + v;
                    xc[h * 4 + v] = j < N ? xxb[j] : 0.0f;
                }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float d[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) d[c] = fmaf(2.0f, acc[r][c], -xq[r]) - xc[c];
                float* row = s.Ds + (ty * 8 + r) * T_COLS;
                *reinterpret_cast<float4*>(row + tx * 4) = make_float4(d[0], d[1], d[2], d[3]);
                *reinterpret_cast<float4*>(row + 64 + tx * 4) = make_float4(d[4], d[5], d[6], d[7]);
            }
            __syncthreads();

            // ---- selection: warp w owns rows w*16 .. w*16+15.  Candidates not below the row's threshold are appended
            // to the row's buffer; the buffer is merged into the sorted list only when it would overflow. ----
            const bool full_tile = j0 + T_COLS <= N;
#pragma unroll 2
            for (int rr = 0; rr < 16; ++rr) {
                const int q = warp * 16 + rr;
                const float4 dv4 = *reinterpret_cast<const float4*>(s.Ds + q * T_COLS + lane * 4);
                const float dv[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
                float tau = s.tau[q];
                const int jb = j0 + lane * 4;
                bool p[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) p[v] = (dv[v] >= tau) && (full_tile || jb + v < N);
                if (__ballot_sync(0xffffffffu, p[0] | p[1] | p[2] | p[3]) == 0u) continue;
                int fill = s.fill[q];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    unsigned m = __ballot_sync(0xffffffffu, p[v]);
                    if (m == 0u) continue;
                    int cnt = __popc(m);
                    if (fill + cnt > 32) {
                        tau = knn_flush(s, q, lane, fill, k);
                        fill = 0;
                        __syncwarp();
#pragma unroll
                        for (int w = v; w < 4; ++w) p[w] = p[w] && (dv[w] >= tau);
                        m = __ballot_sync(0xffffffffu, p[v]);
                        cnt = __popc(m);
                    }
                    if (p[v]) s.buf[q * 32 + fill + __popc(m & lt_mask)] = make_uint2(__float_as_uint(dv[v]), (uint32_t)(jb + v));
                    fill += cnt;
                }
                if (lane == 0) s.fill[q] = fill;
            }
        }
        __syncthreads();
    }

    // ---- final merge and write-out: the k nearest, sorted nearest first ----
    for (int rr = 0; rr < 16; ++rr) {
        const int q = warp * 16 + rr;
        const int fill = s.fill[q];
        if (fill > 0) knn_flush(s, q, lane, fill, k);
        __syncwarp();
        const int n = q0 + q;
        if (n < N && lane < k) {
            const uint2 e = s.list[q * 32 + lane];
            const int64_t o = ((int64_t)b * N + n) * k + lane;
            idx_out[o] = (int32_t)(~e.y);
            if (dist_out) dist_out[o] = ord_val(e.x);
        }
    }
}

}  // namespace gfs

extern "C" int gfs_knn_f32(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, int32_t* idx_out,
                           float* dist_out, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x && sqnorm && idx_out, GFS_ERR_BAD_ARG, "gfs_knn_f32: null pointer");
    GFS_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, GFS_ERR_BAD_ARG, "gfs_knn_f32: non-positive size (B=%d C=%d N=%d k=%d)", B, C, N, k);
    GFS_REQUIRE(k <= N, GFS_ERR_BAD_ARG, "gfs_knn_f32: k=%d exceeds N=%d", k, N);
    GFS_REQUIRE(k <= 32, GFS_ERR_UNSUPPORTED, "gfs_knn_f32: k=%d > 32 is not built (warp-level list holds 32)", k);
    GFS_REQUIRE(C <= 64, GFS_ERR_UNSUPPORTED, "gfs_knn_f32: C=%d > 64 is not built", C);
    GFS_REQUIRE(N % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, GFS_ERR_UNSUPPORTED,
                "gfs_knn_f32: needs N %% 4 == 0 and 16-byte aligned rows (N=%d)", N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sqnorm_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(x, x_bstride, C, N, sqnorm);
    GFS_LAUNCH_OK("sqnorm_kernel");
    const size_t smem = sizeof(KnnSmem) + (size_t)C * T_ROWS * sizeof(float);
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(knn_kernel), sizeof(KnnSmem) + 64 * T_ROWS * sizeof(float)));
    knn_kernel<<<dim3((N + T_ROWS - 1) / T_ROWS, B), T_THREADS, smem, st>>>(x, x_bstride, C, N, k, sqnorm, idx_out, dist_out);
    GFS_LAUNCH_OK("knn_kernel");
    return GFS_OK;
}
