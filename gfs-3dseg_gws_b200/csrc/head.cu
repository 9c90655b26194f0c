// head.cu -- the small fused reductions of the geometric-word head.
//
//   gfs_cos_logits   model/capl.py:290-322 (get_pred: 2 x normalize, proto @ x, x10) fused with
//                    :127-128,:188 (get_gp_weight: weight th where coding[c, assignment] == 1; logits *= weight)
//   gfs_softmax_pool model/capl.py:262-265 (softmax over the POINTS of a block, then pred @ point_feat^T)
#include "common.cuh"

namespace gfs {

constexpr int CL_MAXC = 32;

// One thread per point; prototypes (already L2-normalised) broadcast from shared memory.  All 512 CTAs of the bench shape are
// resident at once, so the kernel lasts as long as ONE thread's chain of dependent load batches: the batches are 32 loads deep
// (4 batches for D = 128 instead of 16 -- ncu r2 of the 8-deep version: 21 % warps active, 13 % of DRAM peak) and the first
// batch is issued before the prototype tile is filled, so its latency hides behind that prologue.
constexpr int CL_BATCH = 32;

// CP = classes padded to 16 / 24 / 32, a template parameter: with a run-time bound the unrolled 32-class loop issues its
// predicated-off fmas and loads too (ncu r2: 52 % issue active for 13 classes).
template <int CP>
__global__ void __launch_bounds__(128)
cos_logits_kernel(const float* __restrict__ feat, int64_t bstride, int D, int N, const float* __restrict__ proto, int PB, int CLS,
                  const float* __restrict__ coding, int G, const int32_t* __restrict__ assignment, float th,
                  float* __restrict__ logits) {
    pdl_enter();
    extern __shared__ __align__(16) float sp[];   // [D][CP]: the CLS prototype values of one channel are CP/4 broadcast LDS.128
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int nn = n < N ? n : N - 1;             // tail threads recompute the last point and store nothing
    const float* f = feat + (int64_t)b * bstride + nn;
    float v[CL_BATCH];
    const bool full0 = D >= CL_BATCH;
    if (full0) {
#pragma unroll
        for (int u = 0; u < CL_BATCH; ++u) v[u] = __ldg(f + (int64_t)u * N);
    }
    const float* pb = proto + (PB > 1 ? (int64_t)b * CLS * D : 0);
    for (int i = threadIdx.x; i < CP * D; i += blockDim.x) {
        const int d = i / CP, c = i - d * CP;
        sp[i] = c < CLS ? __ldg(pb + c * D + d) : 0.0f;
    }
    __syncthreads();
    float acc[CP];
#pragma unroll
    for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
    float nrm = 0.0f;
    int d = 0;
    for (; d + CL_BATCH <= D; d += CL_BATCH) {
        if (d > 0) {
#pragma unroll
            for (int u = 0; u < CL_BATCH; ++u) v[u] = __ldg(f + (int64_t)(d + u) * N);
        }
#pragma unroll
        for (int u = 0; u < CL_BATCH; ++u) {
            nrm = fmaf(v[u], v[u], nrm);
            const float* row = sp + (d + u) * CP;
#pragma unroll
            for (int c = 0; c < CP; c += 4) {
                const float4 w = *reinterpret_cast<const float4*>(row + c);
                acc[c] = fmaf(v[u], w.x, acc[c]);
                acc[c + 1] = fmaf(v[u], w.y, acc[c + 1]);
                acc[c + 2] = fmaf(v[u], w.z, acc[c + 2]);
                acc[c + 3] = fmaf(v[u], w.w, acc[c + 3]);
            }
        }
    }
    for (; d < D; ++d) {
        const float x = __ldg(f + (int64_t)d * N);
        nrm = fmaf(x, x, nrm);
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] = fmaf(x, sp[d * CP + c], acc[c]);
    }
    if (n >= N) return;
    const float inv = 10.0f / fmaxf(sqrtf(nrm), 1e-12f);
    int a = 0;
    if (coding) a = assignment[(int64_t)b * N + n];
#pragma unroll
    for (int c = 0; c < CP; ++c) {
        if (c < CLS) {
            float r = acc[c] * inv;
            if (coding && coding[(int64_t)c * G + a] == 1.0f) r *= th;
            logits[((int64_t)b * CLS + c) * N + n] = r;
        }
    }
}

// max and sum-exp over the points of one (block, class) row
__global__ void __launch_bounds__(256)
softmax_stats_kernel(const float* __restrict__ logits, int N, float* __restrict__ stats) {
    pdl_enter();
    __shared__ float red[8];
    const float* row = logits + (int64_t)blockIdx.x * N;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float mx = -INFINITY;
    for (int n = tid; n < N; n += 256) mx = fmaxf(mx, row[n]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float sum = 0.0f;
    for (int n = tid; n < N; n += 256) sum += expf(row[n] - mx);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += red[w];
        stats[blockIdx.x * 2 + 0] = mx;
        stats[blockIdx.x * 2 + 1] = t;
    }
}

// CTA = (block b, chunk of 128 points): partial[b][chunk][c][d] = sum_{n in chunk} p[c][n] * feat[b][d][n]
constexpr int SP_CH = 64;     // 64-point chunks: 6 CTAs per SM instead of 3 (the kernel is latency bound: load, then reduce)
constexpr int SP_LD = SP_CH + 4;   // row pitch of the staged feature tile: 16-byte aligned rows for 16-byte cp.async
template <int CP>                    // classes padded to 16 / 24 / 32 (see cos_logits_kernel)
__global__ void __launch_bounds__(128)
softmax_pool_kernel(const float* __restrict__ logits, const float* __restrict__ stats, const float* __restrict__ feat,
                    int64_t bstride, int CLS, int D, int N, int nchunks, float* __restrict__ partial) {
    pdl_enter();
    extern __shared__ __align__(16) float sm[];
    float* P = sm;                       // [SP_CH][CP]: the CLS probabilities of one point are CP/4 broadcast LDS.128
    float* F = sm + CP * SP_CH;          // [D][SP_LD]
    const int b = blockIdx.y, ch = blockIdx.x, n0 = ch * SP_CH, tid = threadIdx.x;
    for (int i = tid; i < CP * SP_CH; i += 128) {
        const int c = i / SP_CH, j = i - c * SP_CH;
        const int n = n0 + j;
        float v = 0.0f;
        if (n < N && c < CLS) {
            const float* st = stats + ((int64_t)b * CLS + c) * 2;
            v = expf(logits[((int64_t)b * CLS + c) * N + n] - st[0]) / st[1];
        }
        P[j * CP + c] = v;
    }
    const bool vec = ((N & 3) == 0) && ((bstride & 3) == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0);
    if (vec) {
        for (int i = tid; i < D * (SP_CH / 4); i += 128) {   // 16-byte cp.async: the whole tile is in flight at once
            const int d = i / (SP_CH / 4), j = (i - d * (SP_CH / 4)) * 4;
            const int n = n0 + j;
            const float* src = feat + (int64_t)b * bstride + (int64_t)d * N + (n < N ? n : 0);
            cp_async16(F + d * SP_LD + j, src, n < N ? 16 : 0);
        }
    } else {
        for (int i = tid; i < D * SP_CH; i += 128) {
            const int d = i / SP_CH, j = i - d * SP_CH;
            const int n = n0 + j;
            const float* src = feat + (int64_t)b * bstride + (int64_t)d * N + (n < N ? n : 0);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(smem_u32(F + d * SP_LD + j)), "l"(src),
                         "r"(n < N ? 4 : 0)
                         : "memory");
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int d = tid; d < D; d += 128) {
        float acc[CP];
#pragma unroll
        for (int c = 0; c < CP; ++c) acc[c] = 0.0f;
        const float* fr = F + d * SP_LD;
#pragma unroll 4
        for (int j = 0; j < SP_CH; ++j) {
            const float v = fr[j];
            const float* pr = P + j * CP;
#pragma unroll
            for (int c = 0; c < CP; c += 4) {
                const float4 w = *reinterpret_cast<const float4*>(pr + c);
                acc[c] = fmaf(w.x, v, acc[c]);
                acc[c + 1] = fmaf(w.y, v, acc[c + 1]);
                acc[c + 2] = fmaf(w.z, v, acc[c + 2]);
                acc[c + 3] = fmaf(w.w, v, acc[c + 3]);
            }
        }
#pragma unroll
        for (int c = 0; c < CP; ++c)
            if (c < CLS) partial[(((int64_t)b * nchunks + ch) * CLS + c) * D + d] = acc[c];
    }
}

__global__ void softmax_pool_reduce_kernel(const float* __restrict__ partial, int nchunks, int CD, int64_t total,
                                           float* __restrict__ out) {
    pdl_enter();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / CD, r = i - b * CD;
    float acc = 0.0f;
    for (int ch = 0; ch < nchunks; ++ch) acc += partial[(b * nchunks + ch) * CD + r];   // fixed order: deterministic
    out[i] = acc;
}

// model/capl.py:267-288 (post_refine_proto_v2, eqn. 6) + :117-120 (base classes keep the refined prototype plus the
// generated one, novel classes take the generated one) + the L2 normalisation get_pred applies next, in ONE launch instead
// of ~25 elementwise/reduction kernels on (B, CLS, D) tensors.  One warp per (block, class) row.
//   w = max(<pp/|pp|, q/|q|>, 0);  r = w pp + (1 - w) q;  r = c < base ? r + g : r * 0 + g;  out = r / max(|r|, 1e-12)
__global__ void __launch_bounds__(128)
refine_proto_kernel(const float* __restrict__ pred_proto, const float* __restrict__ proto, const float* __restrict__ gened,
                    int rows, int CLS, int D, int base_num, float* __restrict__ out) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int c = row % CLS;
    const float* pp = pred_proto + (int64_t)row * D;
    const float* q = proto + (int64_t)c * D;
    const float* g = gened + (int64_t)c * D;
    float spp = 0.f, sqq = 0.f, spq = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float a = pp[d], b = q[d];
        spp = fmaf(a, a, spp);
        sqq = fmaf(b, b, sqq);
        spq = fmaf(a, b, spq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        spp += __shfl_xor_sync(0xffffffffu, spp, o);
        sqq += __shfl_xor_sync(0xffffffffu, sqq, o);
        spq += __shfl_xor_sync(0xffffffffu, spq, o);
    }
    float w = spq / (fmaxf(sqrtf(spp), 1e-12f) * fmaxf(sqrtf(sqq), 1e-12f));
    w = w > 0.0f ? w : 0.0f;
    float srr = 0.f;
    for (int d = lane; d < D; d += 32) {
        float r = w * pp[d] + (1.0f - w) * q[d];
        r = c < base_num ? r + g[d] : r * 0.0f + g[d];
        srr = fmaf(r, r, srr);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) srr += __shfl_xor_sync(0xffffffffu, srr, o);
    const float inv = 1.0f / fmaxf(sqrtf(srr), 1e-12f);
    for (int d = lane; d < D; d += 32) {
        float r = w * pp[d] + (1.0f - w) * q[d];
        r = c < base_num ? r + g[d] : r * 0.0f + g[d];
        out[(int64_t)row * D + d] = r * inv;
    }
}

}  // namespace gfs

extern "C" int gfs_cos_logits(const float* feat, int64_t feat_bstride, int B, int D, int N, const float* proto_l2, int PB,
                              int CLS, const float* coding, int G, const int32_t* assignment, float th, float* logits,
                              void* stream) {
    using namespace gfs;
    GFS_REQUIRE(feat && proto_l2 && logits, GFS_ERR_BAD_ARG, "gfs_cos_logits: null pointer");
    GFS_REQUIRE(B > 0 && D > 0 && N > 0 && CLS > 0, GFS_ERR_BAD_ARG, "gfs_cos_logits: non-positive size");
    GFS_REQUIRE(CLS <= CL_MAXC, GFS_ERR_UNSUPPORTED, "gfs_cos_logits: CLS=%d > %d is not built", CLS, CL_MAXC);
    GFS_REQUIRE(PB == 1 || PB == B, GFS_ERR_BAD_ARG, "gfs_cos_logits: PB=%d must be 1 or B=%d", PB, B);
    GFS_REQUIRE(!coding || (assignment && G > 0), GFS_ERR_BAD_ARG, "gfs_cos_logits: coding needs assignment and G");
    const int CP = CLS <= 16 ? 16 : CLS <= 24 ? 24 : 32;
    const size_t smem = (size_t)CP * D * sizeof(float);
    GFS_REQUIRE(smem <= 48 * 1024, GFS_ERR_UNSUPPORTED, "gfs_cos_logits: CLS*D too large");
    const dim3 grid((N + 127) / 128, B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (CP == 16)
        launch_pdl(cos_logits_kernel<16>, grid, dim3(128), smem, st,
        feat, feat_bstride, D, N, proto_l2, PB, CLS, coding, G, assignment, th, logits);
    else if (CP == 24)
        launch_pdl(cos_logits_kernel<24>, grid, dim3(128), smem, st,
        feat, feat_bstride, D, N, proto_l2, PB, CLS, coding, G, assignment, th, logits);
    else
        launch_pdl(cos_logits_kernel<32>, grid, dim3(128), smem, st,
        feat, feat_bstride, D, N, proto_l2, PB, CLS, coding, G, assignment, th, logits);
    GFS_LAUNCH_OK("cos_logits_kernel");
    return GFS_OK;
}

extern "C" int gfs_softmax_pool(const float* logits, const float* feat, int64_t feat_bstride, int B, int CLS, int D, int N,
                                float* stats, float* partial, float* pred_proto, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(logits && feat && stats && partial && pred_proto, GFS_ERR_BAD_ARG, "gfs_softmax_pool: null pointer");
    GFS_REQUIRE(B > 0 && D > 0 && N > 0 && CLS > 0, GFS_ERR_BAD_ARG, "gfs_softmax_pool: non-positive size");
    GFS_REQUIRE(CLS <= CL_MAXC, GFS_ERR_UNSUPPORTED, "gfs_softmax_pool: CLS=%d > %d is not built", CLS, CL_MAXC);
    const int CP = CLS <= 16 ? 16 : CLS <= 24 ? 24 : 32;
    const size_t smem = ((size_t)CP * SP_CH + (size_t)D * SP_LD) * sizeof(float);
    GFS_REQUIRE(smem <= 200 * 1024, GFS_ERR_UNSUPPORTED, "gfs_softmax_pool: D=%d too large", D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nchunks = (N + SP_CH - 1) / SP_CH;
    launch_pdl(softmax_stats_kernel, dim3((unsigned)(B * CLS)), dim3(256), 0, st,
        logits, N, stats);
    GFS_LAUNCH_OK("softmax_stats_kernel");
    const dim3 grid(nchunks, B);
    if (CP == 16) {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(softmax_pool_kernel<16>), 200 * 1024));
        launch_pdl(softmax_pool_kernel<16>, grid, dim3(128), smem, st,
        logits, stats, feat, feat_bstride, CLS, D, N, nchunks, partial);
    } else if (CP == 24) {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(softmax_pool_kernel<24>), 200 * 1024));
        launch_pdl(softmax_pool_kernel<24>, grid, dim3(128), smem, st,
        logits, stats, feat, feat_bstride, CLS, D, N, nchunks, partial);
    } else {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(softmax_pool_kernel<32>), 200 * 1024));
        launch_pdl(softmax_pool_kernel<32>, grid, dim3(128), smem, st,
        logits, stats, feat, feat_bstride, CLS, D, N, nchunks, partial);
    }
    GFS_LAUNCH_OK("softmax_pool_kernel");
    const int64_t total = (int64_t)B * CLS * D;
    launch_pdl(softmax_pool_reduce_kernel, dim3((unsigned)((unsigned)((total + 255) / 256))), dim3(256), 0, st,
        partial, nchunks, CLS * D, total, pred_proto);
    GFS_LAUNCH_OK("softmax_pool_reduce_kernel");
    return GFS_OK;
}

extern "C" int gfs_refine_proto(const float* pred_proto, const float* proto, const float* gened_proto, int B, int CLS, int D,
                                int base_num, float* refine_l2, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(pred_proto && proto && gened_proto && refine_l2, GFS_ERR_BAD_ARG, "gfs_refine_proto: null pointer");
    GFS_REQUIRE(B > 0 && CLS > 0 && D > 0 && base_num >= 0 && base_num <= CLS, GFS_ERR_BAD_ARG,
                "gfs_refine_proto: bad sizes (B=%d CLS=%d D=%d base_num=%d)", B, CLS, D, base_num);
    const int rows = B * CLS;
    launch_pdl(refine_proto_kernel, dim3((unsigned)((rows + 3) / 4)), dim3(128), 0, static_cast<cudaStream_t>(stream),
        pred_proto, proto, gened_proto, rows, CLS, D,
                                                                                      base_num, refine_l2);
    GFS_LAUNCH_OK("refine_proto_kernel");
    return GFS_OK;
}
