// rowsel.cu -- "points x dictionary" fp32 contractions with a fused per-row selection epilogue:
//
//   gfs_gw_project    model/capl.py:344-353  cos = <gp_l2[g], ec>/max(|ec|,1e-12) -> softmax_g(10 cos), argmax_g
//   gfs_kmeans_assign sklearn _k_means_lloyd.pyx:196-218 (as driven by get_basis.py:210): argmin_c (|c|^2 - 2 x.c)
//
// Both are (points x D) . (D x <=192) products evaluated on CUDA cores as fma chains over the channel index ascending
// (the pinned order of oracle/gfs_oracle.c), because the integer result (assignment / label) has to be bit-stable.
// A CTA of 128 threads owns 64 points x 192 dictionary columns; thread (ty, tx) = (tid/16, tid%16) holds 8 points x 12
// columns {h*64 + tx*4 + v}; the 16 threads that share a point row are a half-warp, so the row reduction (max / argmax /
// sum, or argmin) is four xor-shuffles.  Channels stream through shared memory in chunks of 32 (cp.async, 2 stages).
#include "fp32_tile.cuh"

namespace gfs {

constexpr int RS_COLS = 192;
constexpr int RS_KC = 32;

struct RsSmem {
    float As[2][RS_KC * T_ROWS];    // 16 KB
    float Bs[2][RS_KC * RS_COLS];   // 48 KB
};

enum { RS_GW = 0, RS_KMEANS = 1 };

template <int MODE>
__global__ void __launch_bounds__(T_THREADS, 3)
rowsel_kernel(const float* __restrict__ x, int64_t bstride, int D, int N, const float* __restrict__ dict_t, int G, int Gp,
              const float* __restrict__ cnorm, uint8_t* __restrict__ cos_act, int kblocks, int kb0, float* __restrict__ cos_cm,
              int32_t* __restrict__ sel, float* __restrict__ score) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem& s = *reinterpret_cast<RsSmem*>(smem_raw);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.y;
    const int n0 = blockIdx.x * T_ROWS;
    const float* xb = x + (int64_t)b * bstride;
    const int nch = (D + RS_KC - 1) / RS_KC;

    auto load_stage = [&](int ch, int buf) {
        const int c0 = ch * RS_KC;
        const int cn = (D - c0) < RS_KC ? (D - c0) : RS_KC;
        load_panel_async(s.As[buf], T_ROWS, xb + (int64_t)c0 * N, N, cn, n0, N, tid);
        load_panel_async(s.Bs[buf], RS_COLS, dict_t + (int64_t)c0 * Gp, Gp, cn, 0, Gp, tid);
        cp_async_commit();
    };

    float2 acc2[8][6];     // packed FFMA2 accumulators: (column 2j, column 2j+1)
    float nrm[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        nrm[r] = 0.0f;
#pragma unroll
        for (int c = 0; c < 6; ++c) acc2[r][c] = make_float2(0.0f, 0.0f);
    }

    load_stage(0, 0);
    for (int ch = 0; ch < nch; ++ch) {
        if (ch + 1 < nch) {
            load_stage(ch + 1, (ch + 1) & 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const float* As = s.As[ch & 1];
        const float* Bs = s.Bs[ch & 1];
        const int c0 = ch * RS_KC;
        const int cn = (D - c0) < RS_KC ? (D - c0) : RS_KC;
#pragma unroll 2
        for (int c = 0; c < cn; ++c) {
            const float4 a0 = *reinterpret_cast<const float4*>(As + c * T_ROWS + ty * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(As + c * T_ROWS + ty * 8 + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float2 bb[6];
#pragma unroll
            for (int h = 0; h < 3; ++h) {
                const float4 t = *reinterpret_cast<const float4*>(Bs + c * RS_COLS + h * 64 + tx * 4);
                bb[h * 2 + 0] = make_float2(t.x, t.y);
                bb[h * 2 + 1] = make_float2(t.z, t.w);
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (MODE == RS_GW) nrm[r] = fmaf(a[r], a[r], nrm[r]);
                const float2 aa = make_float2(a[r], a[r]);
#pragma unroll
                for (int j = 0; j < 6; ++j) acc2[r][j] = __ffma2_rn(aa, bb[j], acc2[r][j]);
            }
        }
        __syncthreads();
    }
    float acc[8][12];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            acc[r][2 * j] = acc2[r][j].x;
            acc[r][2 * j + 1] = acc2[r][j].y;
        }

    // ------------------------------- epilogue -------------------------------
    int col[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) col[j] = (j >> 2) * 64 + tx * 4 + (j & 3);

    if (MODE == RS_KMEANS) {
        float cn_[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) cn_[j] = col[j] < G ? cnorm[col[j]] : 0.0f;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            float best = INFINITY;
            int bi = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const float sc = fmaf(-2.0f, acc[r][j], cn_[j]);
                // strict '<' with lowest index on ties; columns are visited in increasing index inside each h-group only,
                // so compare (score, index) lexicographically
                if (col[j] < G && (sc < best || (sc == best && col[j] < bi))) {
                    best = sc;
                    bi = col[j];
                }
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob < best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            const int n = n0 + ty * 8 + r;
            if (tx == 0 && n < N) {
                sel[(int64_t)b * N + n] = bi;
                if (score) score[(int64_t)b * N + n] = best;
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float inv = 10.0f / fmaxf(sqrtf(nrm[r]), 1e-12f);
            float lg[12];
            float best = -INFINITY;
            int bi = 0x7fffffff;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                lg[j] = acc[r][j] * inv;
                if (col[j] < G && (lg[j] > best || (lg[j] == best && col[j] < bi))) {
                    best = lg[j];
                    bi = col[j];
                }
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                lg[j] = col[j] < G ? __expf(lg[j] - best) : 0.0f;
                sum += lg[j];
            }
#pragma unroll
            for (int o = 8; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float rs = 1.0f / sum;
            const int n = n0 + ty * 8 + r;
            if (n < N) {
                const int64_t m = (int64_t)b * N + n;
                if (tx == 0) sel[m] = bi;
                if (cos_act) {
                    uint8_t* tile = cos_act + ((m >> 7) * kblocks + kb0) * 16384;
                    const uint32_t row = (uint32_t)(m & 127);
#pragma unroll
                    for (int h = 0; h < 3; ++h) {
                        if (h * 64 < Gp) {
                            uint2 pk;
                            pk.x = pack_bf16x2(lg[h * 4 + 0] * rs, lg[h * 4 + 1] * rs);
                            pk.y = pack_bf16x2(lg[h * 4 + 2] * rs, lg[h * 4 + 3] * rs);
                            *reinterpret_cast<uint2*>(tile + (int64_t)h * 16384 + sw128(row, tx >> 1) + (tx & 1) * 8) = pk;
                        }
                    }
                }
                if (cos_cm) {
#pragma unroll
                    for (int j = 0; j < 12; ++j)
                        if (col[j] < G) cos_cm[((int64_t)b * G + col[j]) * N + n] = lg[j] * rs;
                }
            }
        }
    }
}

}  // namespace gfs

extern "C" int gfs_gw_project(const float* ec, int64_t ec_bstride, int B, int D, int N, const float* gp_l2t, int G, int Gp,
                              void* cosine_act, int kblocks, int kb0, float* cosine_cm, int32_t* assignment, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(ec && gp_l2t && assignment, GFS_ERR_BAD_ARG, "gfs_gw_project: null pointer");
    GFS_REQUIRE(B > 0 && D > 0 && N > 0 && G > 0, GFS_ERR_BAD_ARG, "gfs_gw_project: non-positive size");
    GFS_REQUIRE(G <= Gp && Gp <= RS_COLS && Gp % 64 == 0, GFS_ERR_UNSUPPORTED,
                "gfs_gw_project: G=%d Gp=%d (need G <= Gp <= 192, Gp %% 64 == 0)", G, Gp);
    GFS_REQUIRE(N % 4 == 0 && ec_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(ec) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(gp_l2t) & 15) == 0,
                GFS_ERR_UNSUPPORTED, "gfs_gw_project: needs N %% 4 == 0 and 16-byte aligned pointers");
    if (cosine_act) GFS_REQUIRE(kb0 >= 0 && kb0 + Gp / 64 <= kblocks, GFS_ERR_BAD_ARG, "gfs_gw_project: output blocks out of range");
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(rowsel_kernel<RS_GW>), sizeof(RsSmem)));
    rowsel_kernel<RS_GW><<<dim3((N + T_ROWS - 1) / T_ROWS, B), T_THREADS, sizeof(RsSmem), static_cast<cudaStream_t>(stream)>>>(
        ec, ec_bstride, D, N, gp_l2t, G, Gp, nullptr, static_cast<uint8_t*>(cosine_act), kblocks, kb0, cosine_cm, assignment,
        nullptr);
    GFS_LAUNCH_OK("rowsel_kernel<GW>");
    return GFS_OK;
}

extern "C" int gfs_kmeans_assign(const float* xt, int64_t n, int D, const float* centers_t, int K, int Kp, float* cnorm,
                                 int32_t* labels, float* score, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(xt && centers_t && cnorm && labels, GFS_ERR_BAD_ARG, "gfs_kmeans_assign: null pointer");
    GFS_REQUIRE(n > 0 && D > 0 && K > 0, GFS_ERR_BAD_ARG, "gfs_kmeans_assign: non-positive size");
    GFS_REQUIRE(n < (int64_t)1 << 31, GFS_ERR_UNSUPPORTED, "gfs_kmeans_assign: n=%lld exceeds 2^31 per shard", (long long)n);
    GFS_REQUIRE(K <= Kp && Kp <= RS_COLS && Kp % 4 == 0, GFS_ERR_UNSUPPORTED,
                "gfs_kmeans_assign: K=%d Kp=%d (need K <= Kp <= 192, Kp %% 4 == 0)", K, Kp);
    GFS_REQUIRE(n % 4 == 0 && (reinterpret_cast<uintptr_t>(xt) & 15) == 0 && (reinterpret_cast<uintptr_t>(centers_t) & 15) == 0,
                GFS_ERR_UNSUPPORTED, "gfs_kmeans_assign: needs n %% 4 == 0 (pad the shard) and 16-byte aligned pointers");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sqnorm_kernel<<<dim3((Kp + 255) / 256, 1), 256, 0, st>>>(centers_t, 0, D, Kp, cnorm);
    GFS_LAUNCH_OK("sqnorm_kernel");
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(rowsel_kernel<RS_KMEANS>), sizeof(RsSmem)));
    const int N = (int)n;
    rowsel_kernel<RS_KMEANS><<<dim3((N + T_ROWS - 1) / T_ROWS, 1), T_THREADS, sizeof(RsSmem), st>>>(
        xt, 0, D, N, centers_t, K, Kp, cnorm, nullptr, 0, 0, nullptr, labels, score);
    GFS_LAUNCH_OK("rowsel_kernel<KMEANS>");
    return GFS_OK;
}
