// gemm_tf32.cu -- the training GEMM on the tensor cores: tcgen05.mma kind::tf32, fp32 accumulator in TMEM.
//
// Same contract as gfs_gemm_f32 (gemm_f32.cu), so every 1x1 conv forward, data gradient, weight gradient and the batched
// attention products of the training path (model/dgcnn.py:53-58,63-80, model/attention.py:43-46, model/capl.py:63-65,
// 435-457 under model.train()) move to the tensor pipe without touching their callers:
//
//   C[r, n] = bias[r] + sum_k Aop(k, r) * Bop(k, n)      Aop(k, r) = a_trans ? A[r*lda + k] : A[k*lda + r]   (B alike)
//
// Mapping.  The point index n (contiguous in memory for the channel-major activations) is the UMMA M dimension = the 128
// TMEM lanes, so an epilogue warp stores 32 consecutive floats of one output row per instruction; the r index is the UMMA
// N dimension (NB <= 256 accumulator columns).  Operands are fp32 in HBM in either orientation; the loader warps bring a
// [rows x 32 k] chunk into the canonical K-major SWIZZLE_128B arrangement (32 fp32 = one 128-byte row) with
//   * k contiguous in memory  -> one LDG.128 + one STS.128 per 16-byte chunk,
//   * row contiguous in memory -> four coalesced LDG.32 (lanes along the rows) + one STS.128,
// both conflict-free in shared memory, rounding to tf32 with round-to-nearest (cvt.rna) on the way -- the tensor core
// would otherwise TRUNCATE the 13 low mantissa bits, a bias of 2^-11 per operand that does not average out over k.
// Several CTAs are resident per SM (48-96 KB of shared memory, NB TMEM columns each), which is what keeps enough loads in
// flight for these HBM-bound shapes; inside a CTA the asynchronous MMAs of chunk c overlap the loads of chunk c+1
// (two shared-memory stages, tcgen05.commit -> mbarrier hands a stage back).
//
// Accuracy: |C - exact| <= 2^-10 * sum_k |Aop||Bop| + fp32 accumulation (tests/test_gpu_train.py::test_gemm_tf32_*).
// That is NOT enough for the backward pass: a z error of 3e-4 flips the LeakyReLU / max-over-k branch of ~3e-4 of the
// elements, which shows up as 1-6 % relative L2 error of the gradients (measured).  The training path therefore runs the
// 3xTF32 mode (split3): hi/lo tf32 pairs, three MMAs per k-step, fp32-grade products (<= ~1e-6), still HBM-bound.
#include "common.cuh"

namespace gfs {

constexpr int TG_THREADS = 256;
constexpr int TG_STAGES = 2;
constexpr int TG_KC = 32;          // fp32 per 128-byte swizzle row

__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// one 16-byte chunk into the swizzled tile: the tf32 rounding of v and, for the 3xTF32 mode, the rounded residual v - hi
// (exact in fp32) into the twin tile lo_off bytes further on
template <bool SPLIT>
__device__ __forceinline__ void tg_store(uint8_t* tile, int lo_off, int row, int q, const float4& v) {
    const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
    uint8_t* p = tile + sw128(row, q);
    *reinterpret_cast<float4*>(p) = h;
    if (SPLIT)
        *reinterpret_cast<float4*>(p + lo_off) =
            make_float4(rna_tf32(v.x - h.x), rna_tf32(v.y - h.y), rna_tf32(v.z - h.z), rna_tf32(v.w - h.w));
}

// tile[row][k] (row < nrows, k < 32) <- op(k0 + k, row0 + row), zero outside [0, kend) x [0, lim)
template <bool SPLIT>
__device__ __forceinline__ void tg_load(uint8_t* tile, int lo_off, int nrows, const float* __restrict__ src, int64_t ld, int trans,
                                        int k0, int kend, int row0, int lim, int tid) {
    const int items = nrows * 8;
    if (trans) {
        // src[row*ld + k]: k contiguous
        const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((k0 & 3) == 0);
        for (int base = tid; base < items; base += TG_THREADS * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * TG_THREADS;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < items) {
                    const int row = i >> 3, q = i & 7;
                    const int k = k0 + q * 4, rr = row0 + row;
                    if (rr < lim && k < kend) {
                        const float* p = src + (int64_t)rr * ld + k;
                        if (vec && k + 3 < kend) {
                            v[u] = __ldg(reinterpret_cast<const float4*>(p));
                        } else {
                            v[u].x = __ldg(p);
                            if (k + 1 < kend) v[u].y = __ldg(p + 1);
                            if (k + 2 < kend) v[u].z = __ldg(p + 2);
                            if (k + 3 < kend) v[u].w = __ldg(p + 3);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * TG_THREADS;
                if (i < items) {
                    const int row = i >> 3, q = i & 7;
                    tg_store<SPLIT>(tile, lo_off, row, q, v[u]);
                }
            }
        }
    } else {
        // src[k*ld + row]: rows contiguous -> lanes along the rows
        for (int base = tid; base < items; base += TG_THREADS * 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * TG_THREADS;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < items) {
                    const int q = i / nrows, row = i - q * nrows;
                    const int k = k0 + q * 4, rr = row0 + row;
                    if (rr < lim && k < kend) {
                        const float* p = src + (int64_t)k * ld + rr;
                        v[u].x = __ldg(p);
                        if (k + 1 < kend) v[u].y = __ldg(p + ld);
                        if (k + 2 < kend) v[u].z = __ldg(p + 2 * ld);
                        if (k + 3 < kend) v[u].w = __ldg(p + 3 * ld);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * TG_THREADS;
                if (i < items) {
                    const int q = i / nrows, row = i - q * nrows;
                    tg_store<SPLIT>(tile, lo_off, row, q, v[u]);
                }
            }
        }
    }
}

// grid: (ceil(Ncols/128), ceil(R/NB), batch*splitk); dynamic smem: 1024 (alignment) + stages*(SPLIT ? 2 : 1)*(16384 + NB*128)
// SPLIT = 3xTF32: each operand is held as hi + lo (two tf32 tiles) and every k-step issues hi*hi + hi*lo + lo*hi, which
// recovers fp32-grade products (the dropped lo*lo term and the rounding of lo are both ~2^-22 relative) on the tensor pipe.
template <bool SPLIT>
__global__ void __launch_bounds__(TG_THREADS, 3)
gemm_tf32_kernel(const float* __restrict__ A, int64_t lda, int a_trans, int64_t a_bs, const float* __restrict__ B, int64_t ldb,
                 int b_trans, int64_t b_bs, float* __restrict__ C, int64_t ldc, int c_trans, int64_t c_bs,
                 const float* __restrict__ bias, int R, int Ncols, int K, int splitk, int kper, int NB, int tmem_cols, int stages) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_empty[TG_STAGES];
    __shared__ uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int lo_off = 16384 + NB * 128;                          // hi tiles of both operands, then their lo twins
    const int stage_bytes = (SPLIT ? 2 : 1) * lo_off;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * 128, r0 = blockIdx.y * NB;
    const int z = blockIdx.z;
    const int bz = z / splitk, sk = z - bz * splitk;
    A += (int64_t)bz * a_bs;
    B += (int64_t)bz * b_bs;
    const int kbeg = sk * kper;
    const int kend = (kbeg + kper) < K ? (kbeg + kper) : K;
    const int nk = kend > kbeg ? (kend - kbeg + TG_KC - 1) / TG_KC : 0;

    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < TG_STAGES; ++s) mbar_init(&bar_empty[s], 1);
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = umma_idesc_tf32(128, (uint32_t)NB);

    for (int c = 0; c < nk; ++c) {
        const int s = c % stages;
        uint8_t* tl = smem + s * stage_bytes;        // [128 n rows][32 k]  = UMMA A
        uint8_t* tc = tl + 16384;                    // [NB r rows][32 k]   = UMMA B
        if (c >= stages) mbar_wait(&bar_empty[s], (uint32_t)((c / stages - 1) & 1));
        const int k0 = kbeg + c * TG_KC;
        tg_load<SPLIT>(tl, lo_off, 128, B, ldb, b_trans, k0, kend, n0, Ncols, tid);
        tg_load<SPLIT>(tc, lo_off, NB, A, lda, a_trans, k0, kend, r0, R, tid);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t ad = umma_desc_sw128(smem_u32(tl)), bd = umma_desc_sw128(smem_u32(tc));
            const uint64_t lo = (uint64_t)(lo_off >> 4);
            const int kn = (kend - k0) < TG_KC ? (kend - k0) : TG_KC;
            for (int kk = 0; kk * 8 < kn; ++kk) {
                const uint64_t a = ad + (uint64_t)(kk * 2), b = bd + (uint64_t)(kk * 2);
                if (SPLIT) {                          // the two small cross terms first, then the leading term
                    umma_tf32(tmem, a + lo, b, idesc, (c > 0 || kk > 0) ? 1u : 0u);
                    umma_tf32(tmem, a, b + lo, idesc, 1u);
                    umma_tf32(tmem, a, b, idesc, 1u);
                } else {
                    umma_tf32(tmem, a, b, idesc, (c > 0 || kk > 0) ? 1u : 0u);
                }
            }
            umma_commit(&bar_empty[s]);
            if (c == nk - 1) umma_commit(&bar_done);
        }
    }
    if (nk > 0) {
        mbar_wait(&bar_done, 0);
        tc_fence_after();
    }

    // epilogue: warp w owns TMEM lanes 32*(w&3).. and every other 32-column block
    float* Cb = C + (splitk > 1 ? (int64_t)z * R * Ncols : (int64_t)bz * c_bs);
    const bool row_major = splitk > 1 || !c_trans;
    const int64_t ld = splitk > 1 ? Ncols : ldc;
    const int n = n0 + (warp & 3) * 32 + lane;
    const bool use_bias = bias != nullptr && splitk == 1;
    for (int cb = (warp >> 2); cb * 32 < NB; cb += 2) {
        uint32_t v[32];
        if (nk > 0) {
            tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(cb * 32), v);
            tmem_ld_wait32(v);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        const int rb = r0 + cb * 32;
        if (n < Ncols) {
            if (row_major) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int r = rb + j;
                    if (r < R && cb * 32 + j < NB) Cb[(int64_t)r * ld + n] = __uint_as_float(v[j]) + (use_bias ? __ldg(bias + r) : 0.0f);
                }
            } else {
                float* o = Cb + (int64_t)n * ld + rb;
                const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cb) & 15) == 0) && ((rb & 3) == 0);
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (vec && rb + j + 3 < R && cb * 32 + j + 3 < NB) {
                        float4 t;
                        t.x = __uint_as_float(v[j]) + (use_bias ? __ldg(bias + rb + j) : 0.0f);
                        t.y = __uint_as_float(v[j + 1]) + (use_bias ? __ldg(bias + rb + j + 1) : 0.0f);
                        t.z = __uint_as_float(v[j + 2]) + (use_bias ? __ldg(bias + rb + j + 2) : 0.0f);
                        t.w = __uint_as_float(v[j + 3]) + (use_bias ? __ldg(bias + rb + j + 3) : 0.0f);
                        *reinterpret_cast<float4*>(o + j) = t;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (rb + j + e < R && cb * 32 + j + e < NB)
                                o[j + e] = __uint_as_float(v[j + e]) + (use_bias ? __ldg(bias + rb + j + e) : 0.0f);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

__global__ void gemm_tf32_reduce_kernel(const float* __restrict__ part, int splitk, int R, int Ncols, const float* __restrict__ bias,
                                        float* __restrict__ C, int64_t ldc, int c_trans, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R * Ncols) return;
    const int r = i / Ncols, n = i - r * Ncols;
    float acc = bias ? bias[r] : 0.0f;
    for (int s = 0; s < splitk; ++s) acc += part[(int64_t)s * R * Ncols + i];   // fixed order: deterministic
    float* o = c_trans ? C + (int64_t)n * ldc + r : C + (int64_t)r * ldc + n;
    *o = accumulate ? *o + acc : acc;
}

}  // namespace gfs

extern "C" int gfs_gemm_tf32(const float* A, int64_t lda, int a_trans, int64_t a_bstride, const float* B, int64_t ldb, int b_trans,
                             int64_t b_bstride, float* C, int64_t ldc, int c_trans, int64_t c_bstride, const float* bias, int R,
                             int Ncols, int K, int batch, int splitk, float* workspace, int accumulate, int split3, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(A && B && C, GFS_ERR_BAD_ARG, "gfs_gemm_tf32: null pointer");
    GFS_REQUIRE(R > 0 && Ncols > 0 && K > 0 && batch > 0 && splitk > 0, GFS_ERR_BAD_ARG, "gfs_gemm_tf32: non-positive size");
    GFS_REQUIRE(splitk == 1 || (batch == 1 && workspace), GFS_ERR_BAD_ARG, "gfs_gemm_tf32: split-K needs batch == 1 and a workspace");
    GFS_REQUIRE(splitk > 1 || !accumulate, GFS_ERR_UNSUPPORTED, "gfs_gemm_tf32: accumulate is only built for the split-K reduction");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int kper = (K + splitk - 1) / splitk;
    kper = (kper + TG_KC - 1) / TG_KC * TG_KC;
    // accumulator columns per CTA: R in equal tiles of at most 256, rounded up to the UMMA N granularity of 16 (M = 128)
    const int rt = (R + 255) / 256;
    int NB = ((R + rt - 1) / rt + 15) / 16 * 16;
    int tmem_cols = 32;
    while (tmem_cols < NB) tmem_cols *= 2;
    const dim3 grid((Ncols + 127) / 128, (R + NB - 1) / NB, batch * splitk);
    GFS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, GFS_ERR_UNSUPPORTED, "gfs_gemm_tf32: grid too large");
    // Resident CTAs per SM are what keeps enough loads in flight for these HBM-bound shapes (measured: 4 CTAs x 1 stage beat
    // 2 CTAs x 2 stages by 15 % over the training step), so the second shared-memory stage -- MMAs of chunk c overlapping the
    // loads of chunk c+1 inside one CTA -- is only taken when it costs no residency (limits: 4 CTAs by registers, 512 TMEM columns)
    const size_t stage = (size_t)(split3 ? 2 : 1) * (16384 + (size_t)NB * 128);
    auto resident = [&](int st) {
        const int by_smem = (int)((227 * 1024) / (1024 + st * stage + 1024));
        const int by_tmem = 512 / tmem_cols;
        return by_smem < by_tmem ? (by_smem < 4 ? by_smem : 4) : (by_tmem < 4 ? by_tmem : 4);
    };
    const int stages = resident(2) == resident(1) ? 2 : 1;
    const size_t smem = 1024 + stages * stage;
    float* out = splitk > 1 ? workspace : C;
    if (split3) {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(gemm_tf32_kernel<true>), 1024 + 2 * (16384 + 256 * 128)));
        gemm_tf32_kernel<true><<<grid, TG_THREADS, smem, st>>>(A, lda, a_trans, a_bstride, B, ldb, b_trans, b_bstride, out, ldc, c_trans,
                                                               c_bstride, bias, R, Ncols, K, splitk, kper, NB, tmem_cols, stages);
    } else {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(gemm_tf32_kernel<false>), 1024 + 2 * (16384 + 256 * 128)));
        gemm_tf32_kernel<false><<<grid, TG_THREADS, smem, st>>>(A, lda, a_trans, a_bstride, B, ldb, b_trans, b_bstride, out, ldc, c_trans,
                                                                c_bstride, bias, R, Ncols, K, splitk, kper, NB, tmem_cols, stages);
    }
    GFS_LAUNCH_OK("gemm_tf32_kernel");
    if (splitk > 1) {
        gemm_tf32_reduce_kernel<<<(R * Ncols + 255) / 256, 256, 0, st>>>(workspace, splitk, R, Ncols, bias, C, ldc, c_trans, accumulate);
        GFS_LAUNCH_OK("gemm_tf32_reduce_kernel");
    }
    return GFS_OK;
}
