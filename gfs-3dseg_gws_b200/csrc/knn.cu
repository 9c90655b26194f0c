// knn.cu -- fused fp32 kNN graph (replaces model/dgcnn.py:17-23: matmul + 4 elementwise passes + topk).
//
// One CTA owns 64 query points of one block and streams all N candidates in tiles of 128 through shared memory
// (cp.async, double buffered).  The CTA is warp-specialised: warps 0-3 form distance tiles with the shared FFMA core,
// warps 4-7 run the top-k selection on the previous tile (two distance tiles in flight, mbarrier hand-off), so the
// FMA pipe and the ALU/shuffle-bound selection overlap.  Distances are formed in registers in the pinned order
//     dot = fma chain over c ascending;  d = fmaf(2, dot, -xx_i) - xx_j
// and dropped into a 64x128 shared tile; each warp then scans its 16 rows against the row's threshold (the k-th best at
// the last merge: a one-compare filter that rejects ~97 % of candidates after the first tile) and appends survivors to a
// 32-entry per-row buffer.  Only when the buffer would overflow (about 4 times per row at N=2048) is it merged into the
// row's sorted list: a warp-wide bitonic sort of 64-bit keys (orderable distance << 32 | ~index) followed by a bitonic
// top-32 merge -- 20 compare-exchange stages, no serial insertion.  Order: larger d first, ties -> smaller index.
// The N x N matrix never leaves the SM.
#include "fp32_tile.cuh"

namespace gfs {

constexpr int KNN_KC = 64;        // channels per pipeline stage (all of C <= 64: one stage per candidate tile)
constexpr int KNN_SEL_WARPS = 16;
constexpr int KNN_ROWS_PER_SEL = T_ROWS / KNN_SEL_WARPS;
constexpr int KNN_THREADS = 128 + 32 * KNN_SEL_WARPS;  // warps 0-3: FFMA producers of distance tiles; the rest: selection consumers

struct KnnSmem {
    float Bs[2][KNN_KC * T_COLS];   // 32 KB  candidate panels (cp.async double buffer)
    float Ds[2][T_ROWS * T_COLS];   // 64 KB  distance tiles, produced by the FFMA warps, consumed by the selection warps
    uint2 list[T_ROWS * 64];        // 32 KB  per query: best 32 (k <= 32) or 64 (k <= 64) so far, sorted descending;
                                    //        (x = orderable distance, y = ~index); rank r lives at [q*64 + r]
    uint2 buf[T_ROWS * 32];         // 16 KB  per query: raw (distance bits, index) appended since the last merge
    float tau[T_ROWS];              // distance of the k-th best at the last merge (filter threshold, never decreases)
    int fill[T_ROWS];
    uint64_t ds_full[2], ds_empty[2];
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
// (__noinline__: the merge is ~400 instructions and is reached from several places of the selection loop; one shared copy
// keeps the kernel's code footprint small, the FFMA warps' loop competes for the same instruction cache)
template <int LR>   // LR = list registers per lane: 1 -> 32-entry list (k <= 32), 2 -> 64-entry list (k <= 64)
__device__ __noinline__ float knn_flush(KnnSmem& s, int q, int lane, int fill, int k) {
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
    const u64 rev = __shfl_sync(0xffffffffu, key, 31 - lane);
    uint32_t kth;
    if (LR == 1) {
        const uint2 l2 = s.list[q * 64 + lane];
        const u64 lst = ((u64)l2.x << 32) | (u64)l2.y;
        key = lst > rev ? lst : rev;
#pragma unroll
        for (int j2 = 16; j2 > 0; j2 >>= 1) key = cmpex(key, j2, (lane & j2) == 0);
        s.list[q * 64 + lane] = make_uint2((uint32_t)(key >> 32), (uint32_t)key);
        kth = __shfl_sync(0xffffffffu, (uint32_t)(key >> 32), k - 1);
    } else {
        // 64-entry list: ranks 0..31 in m0, 32..63 in m1.  union top-64 = bitonic merge of [L0, max(L1, reverse(S))]
        const uint2 a2 = s.list[q * 64 + lane], b2 = s.list[q * 64 + 32 + lane];
        u64 m0 = ((u64)a2.x << 32) | (u64)a2.y;
        u64 m1 = ((u64)b2.x << 32) | (u64)b2.y;
        m1 = m1 > rev ? m1 : rev;
        const u64 hi = m0 > m1 ? m0 : m1, lo = m0 > m1 ? m1 : m0;
        m0 = hi;
        m1 = lo;
#pragma unroll
        for (int j2 = 16; j2 > 0; j2 >>= 1) {
            m0 = cmpex(m0, j2, (lane & j2) == 0);
            m1 = cmpex(m1, j2, (lane & j2) == 0);
        }
        s.list[q * 64 + lane] = make_uint2((uint32_t)(m0 >> 32), (uint32_t)m0);
        s.list[q * 64 + 32 + lane] = make_uint2((uint32_t)(m1 >> 32), (uint32_t)m1);
        kth = k <= 32 ? __shfl_sync(0xffffffffu, (uint32_t)(m0 >> 32), k - 1)
                      : __shfl_sync(0xffffffffu, (uint32_t)(m1 >> 32), k - 33);
    }
    const float tau = kth ? ord_val(kth) : -INFINITY;
    if (lane == 0) s.tau[q] = tau;
    return tau;
}

template <int LR>
__global__ void __launch_bounds__(KNN_THREADS, 1)
knn_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, int k, const float* __restrict__ sqnorm,
           int32_t* __restrict__ idx_out, float* __restrict__ dist_out, const int* __restrict__ tile_flags) {
    pdl_wait();
    // repair pass of knn_tc.cu: only the flagged 64-row tiles are redone
    if (tile_flags && tile_flags[blockIdx.y * gridDim.x + blockIdx.x] == 0) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KnnSmem& s = *reinterpret_cast<KnnSmem*>(smem_raw);
    float* As = reinterpret_cast<float*>(smem_raw + sizeof(KnnSmem));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // FFMA warps are the LAST four warps: the warp scheduler's arbitration favours higher warp ids, and the FFMA warps
    // are the critical path (the selection warps mostly wait)
    const bool is_fma = warp >= KNN_SEL_WARPS;
    const int tid = is_fma ? (int)threadIdx.x - 32 * KNN_SEL_WARPS : (int)threadIdx.x;
    const int b = blockIdx.y, q0 = blockIdx.x * T_ROWS;
    const int ntiles = (N + T_COLS - 1) / T_COLS;

    for (int i = threadIdx.x; i < T_ROWS * 64; i += KNN_THREADS) s.list[i] = make_uint2(0u, 0u);
    if (threadIdx.x < T_ROWS) {
        s.tau[threadIdx.x] = -INFINITY;
        s.fill[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0) {
        mbar_init(&s.ds_full[0], 128);
        mbar_init(&s.ds_full[1], 128);
        mbar_init(&s.ds_empty[0], 32 * KNN_SEL_WARPS);
        mbar_init(&s.ds_empty[1], 32 * KNN_SEL_WARPS);
        mbar_fence_init();
    }
    __syncthreads();

    if (is_fma) {
        // ======================= FFMA warps: distance tiles in the pinned order =======================
        reg_alloc<168>();
        const int ty = tid >> 4, tx = tid & 15;
        const float* xb = x + (int64_t)b * bstride;
        const float* xxb = sqnorm + (int64_t)b * N;
        const int nch = (C + KNN_KC - 1) / KNN_KC;
        const int S = ntiles * nch;

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
            named_bar_sync(1, 128);

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
                const int db = t & 1;
                mbar_wait_backoff(&s.ds_empty[db], ((t >> 1) & 1) ^ 1, 64);
                float* D = s.Ds[db];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    float d[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) d[c] = fmaf(2.0f, acc[r][c], -xq[r]) - xc[c];
                    float* row = D + (ty * 8 + r) * T_COLS;
                    *reinterpret_cast<float4*>(row + tx * 4) = make_float4(d[0], d[1], d[2], d[3]);
                    *reinterpret_cast<float4*>(row + 64 + tx * 4) = make_float4(d[4], d[5], d[6], d[7]);
                }
                mbar_arrive(&s.ds_full[db]);
            }
            named_bar_sync(1, 128);   // everyone is done with Bs[st & 1] before the next prefetch overwrites it
        }
    } else {
        // ======================= selection warps: threshold filter, append, merge on overflow =======================
        reg_dealloc<72>();
        const int w = warp;
        const unsigned lt_mask = (1u << lane) - 1u;
        for (int t = 0; t < ntiles; ++t) {
            const int db = t & 1;
            const int j0 = t * T_COLS;
            const bool full_tile = j0 + T_COLS <= N;
            mbar_wait_backoff(&s.ds_full[db], (t >> 1) & 1, 256);
            const float* D = s.Ds[db];
            for (int rr = 0; rr < KNN_ROWS_PER_SEL; ++rr) {
                const int q = w * KNN_ROWS_PER_SEL + rr;
                const float4 dv4 = *reinterpret_cast<const float4*>(D + q * T_COLS + lane * 4);
                const float dv[4] = {dv4.x, dv4.y, dv4.z, dv4.w};
                float tau = s.tau[q];
                const int jb = j0 + lane * 4;
                if (LR == 1 && t == 0) {
                    // First tile: every candidate is a survivor (tau = -inf).  Instead of four overflow merges, sort the four
                    // 32-candidate column groups as four interleaved bitonic networks (ILP 4) and merge them pairwise:
                    // the row's list starts as the exact top 32 of its first 128 candidates.
                    u64 key[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        key[v] = (jb + v < N) ? (((u64)ord_key(dv[v]) << 32) | (u64)(~(uint32_t)(jb + v))) : 0ull;
#pragma unroll
                    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
                        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                            const bool keep_max = ((lane & j2) == 0) == ((lane & k2) == 0);
#pragma unroll
                            for (int v = 0; v < 4; ++v) key[v] = cmpex(key[v], j2, keep_max);
                        }
                    }
                    u64 r1 = __shfl_sync(0xffffffffu, key[1], 31 - lane), r3 = __shfl_sync(0xffffffffu, key[3], 31 - lane);
                    u64 t0 = key[0] > r1 ? key[0] : r1, t1 = key[2] > r3 ? key[2] : r3;
#pragma unroll
                    for (int j2 = 16; j2 > 0; j2 >>= 1) {
                        t0 = cmpex(t0, j2, (lane & j2) == 0);
                        t1 = cmpex(t1, j2, (lane & j2) == 0);
                    }
                    const u64 rr1 = __shfl_sync(0xffffffffu, t1, 31 - lane);
                    u64 f = t0 > rr1 ? t0 : rr1;
#pragma unroll
                    for (int j2 = 16; j2 > 0; j2 >>= 1) f = cmpex(f, j2, (lane & j2) == 0);
                    s.list[q * 64 + lane] = make_uint2((uint32_t)(f >> 32), (uint32_t)f);
                    const uint32_t kth = __shfl_sync(0xffffffffu, (uint32_t)(f >> 32), k - 1);
                    if (lane == 0) s.tau[q] = kth ? ord_val(kth) : -INFINITY;
                    continue;
                }
                bool p[4];
                unsigned m[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    p[v] = (dv[v] >= tau) && (full_tile || jb + v < N);
                    m[v] = __ballot_sync(0xffffffffu, p[v]);
                }
                if ((m[0] | m[1] | m[2] | m[3]) == 0u) continue;
                int fill = s.fill[q];
                const int c0 = __popc(m[0]), c1 = __popc(m[1]), c2 = __popc(m[2]), c3 = __popc(m[3]);
                if (fill + c0 + c1 + c2 + c3 <= 32) {
                    // common case: everything fits, one compaction per column group
                    uint2* B = s.buf + q * 32 + fill;
                    if (p[0]) B[__popc(m[0] & lt_mask)] = make_uint2(__float_as_uint(dv[0]), (uint32_t)jb);
                    if (p[1]) B[c0 + __popc(m[1] & lt_mask)] = make_uint2(__float_as_uint(dv[1]), (uint32_t)(jb + 1));
                    if (p[2]) B[c0 + c1 + __popc(m[2] & lt_mask)] = make_uint2(__float_as_uint(dv[2]), (uint32_t)(jb + 2));
                    if (p[3]) B[c0 + c1 + c2 + __popc(m[3] & lt_mask)] = make_uint2(__float_as_uint(dv[3]), (uint32_t)(jb + 3));
                    fill += c0 + c1 + c2 + c3;
                } else {
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        unsigned mm = __ballot_sync(0xffffffffu, p[v]);
                        if (mm == 0u) continue;
                        int cnt = __popc(mm);
                        if (fill + cnt > 32) {
                            tau = knn_flush<LR>(s, q, lane, fill, k);
                            fill = 0;
                            __syncwarp();
#pragma unroll
                            for (int u = v; u < 4; ++u) p[u] = p[u] && (dv[u] >= tau);
                            mm = __ballot_sync(0xffffffffu, p[v]);
                            cnt = __popc(mm);
                        }
                        if (p[v]) s.buf[q * 32 + fill + __popc(mm & lt_mask)] = make_uint2(__float_as_uint(dv[v]), (uint32_t)(jb + v));
                        fill += cnt;
                    }
                }
                if (lane == 0) s.fill[q] = fill;
            }
            __syncwarp();
            mbar_arrive(&s.ds_empty[db]);
        }

        // ---- final merge and write-out: the k nearest, sorted nearest first ----
        for (int rr = 0; rr < KNN_ROWS_PER_SEL; ++rr) {
            const int q = w * KNN_ROWS_PER_SEL + rr;
            const int fill = s.fill[q];
            if (fill > 0) knn_flush<LR>(s, q, lane, fill, k);
            __syncwarp();
            const int n = q0 + q;
            if (n < N) {
#pragma unroll
                for (int h = 0; h < LR; ++h) {
                    const int r = h * 32 + lane;
                    if (r < k) {
                        const uint2 e = s.list[q * 64 + r];
                        const int64_t o = ((int64_t)b * N + n) * k + r;
                        idx_out[o] = (int32_t)(~e.y);
                        if (dist_out) dist_out[o] = ord_val(e.x);
                    }
                }
            }
        }
    }
}

static int knn_exact_launch(const float* x, int64_t x_bstride, int B, int C, int N, int k, const float* sqnorm, const int* flags,
                            int32_t* idx_out, float* dist_out, cudaStream_t st) {
    const size_t smem = sizeof(KnnSmem) + (size_t)C * T_ROWS * sizeof(float);
    const dim3 grid((N + T_ROWS - 1) / T_ROWS, B);
    if (k <= 32) {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(knn_kernel<1>), sizeof(KnnSmem) + 64 * T_ROWS * sizeof(float)));
        launch_pdl<2>(knn_kernel<1>, grid, dim3(KNN_THREADS), smem, st,
        x, x_bstride, C, N, k, sqnorm, idx_out, dist_out, flags);
    } else {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(knn_kernel<2>), sizeof(KnnSmem) + 64 * T_ROWS * sizeof(float)));
        launch_pdl<2>(knn_kernel<2>, grid, dim3(KNN_THREADS), smem, st,
        x, x_bstride, C, N, k, sqnorm, idx_out, dist_out, flags);
    }
    GFS_LAUNCH_OK("knn_kernel");
    return GFS_OK;
}

int knn_exact_flagged(const float* x, int64_t x_bstride, int B, int C, int N, int k, const float* sqnorm, const int* flags,
                      int32_t* idx_out, float* dist_out, cudaStream_t st) {
    return knn_exact_launch(x, x_bstride, B, C, N, k, sqnorm, flags, idx_out, dist_out, st);
}

}  // namespace gfs

extern "C" int gfs_knn_f32(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, int32_t* idx_out,
                           float* dist_out, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x && sqnorm && idx_out, GFS_ERR_BAD_ARG, "gfs_knn_f32: null pointer");
    GFS_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, GFS_ERR_BAD_ARG, "gfs_knn_f32: non-positive size (B=%d C=%d N=%d k=%d)", B, C, N, k);
    GFS_REQUIRE(k <= N, GFS_ERR_BAD_ARG, "gfs_knn_f32: k=%d exceeds N=%d", k, N);
    GFS_REQUIRE(k <= 64, GFS_ERR_UNSUPPORTED, "gfs_knn_f32: k=%d > 64 is not built (the warp-level list holds 64)", k);
    GFS_REQUIRE(C <= 64, GFS_ERR_UNSUPPORTED, "gfs_knn_f32: C=%d > 64 is not built", C);
    GFS_REQUIRE(N % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, GFS_ERR_UNSUPPORTED,
                "gfs_knn_f32: needs N %% 4 == 0 and 16-byte aligned rows (N=%d)", N);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sqnorm_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(x, x_bstride, C, N, sqnorm);
    GFS_LAUNCH_OK("sqnorm_kernel");
    return knn_exact_launch(x, x_bstride, B, C, N, k, sqnorm, nullptr, idx_out, dist_out, st);
}
