// attention.cu -- flash-style single-head self-attention on tcgen05.
//
// Replaces model/attention.py:43-46: attn = softmax(q^T k / sqrt(d)) (an N x N fp32 matrix per block, 537 MB at B=32),
// y = attn . v^T.  Here the N x N matrix never exists: per 128-query tile the scores of one 128-key tile live in TMEM,
// the online-softmax weights go to shared memory as a bf16 UMMA operand, and P.V is a second tcgen05.mma.
//
// Inputs come straight from the fused q/k/v linear layer: one bf16 "act" matrix with three 64-column blocks [q | k | v],
// so every operand tile is a single 16 KiB TMA bulk copy:
//   S = Q K^T : A = Q tile (128 x 64, K-major),  B = K tile (128 keys x 64, K-major)            -> 128 x 128 fp32 in TMEM
//   O = P V   : A = P tile (128 x 128 keys, K-major, written by the softmax warps),
//               B = V tile (128 keys x 64) used as an MN-major operand (d contiguous)             -> 128 x 64 fp32 in TMEM
// Persistent CTA, 6 warps: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-5 softmax/epilogue (one query
// row per thread; running max / sum / output accumulator in registers, so TMEM is never read-modify-written).
#include "common.cuh"

namespace gfs {

constexpr int AT_KV = 3;          // K/V pipeline stages
constexpr int AT_THREADS = 192;

struct AtSmem {
    uint8_t Q[16384];
    uint8_t K[AT_KV][16384];
    uint8_t V[AT_KV][16384];
    uint8_t P[2][32768];          // two key-blocks of 64 per buffer
    uint64_t q_full, q_empty;
    uint64_t k_full[AT_KV], k_empty[AT_KV], v_full[AT_KV], v_empty[AT_KV];
    uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2], o_full[2], o_empty[2];
    uint32_t tmem_base;
};

// K-major SW128 descriptor is umma_desc_sw128(); MN-major SW128: 64 MN elements (128 B) contiguous, 8 K rows per 1024 B atom
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(1024u >> 4) << 16;   // leading byte offset: next 64-wide MN block (unused for N = 64)
    d |= (uint64_t)(1024u >> 4) << 32;   // stride byte offset: next group of 8 K rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(uint32_t M, uint32_t N) {   // B operand MN-major
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__global__ void __launch_bounds__(AT_THREADS, 1)
attention_kernel(const uint8_t* __restrict__ qkv, int kblocks, int kb_q, int tiles_per_block, int n_items, float scale_log2,
                 int N, float* __restrict__ y_cm, int64_t y_bstride, uint8_t* __restrict__ y_act, int y_kblocks, int y_kb) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    // align inside the shared window with pointer arithmetic on smem_raw (keeps the .shared address space: LDS/STS, not generic LD/ST)
    AtSmem& s = *reinterpret_cast<AtSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = tiles_per_block;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&s.q_full, 1);
            mbar_init(&s.q_empty, 1);
            for (int i = 0; i < AT_KV; ++i) {
                mbar_init(&s.k_full[i], 1);
                mbar_init(&s.k_empty[i], 1);
                mbar_init(&s.v_full[i], 1);
                mbar_init(&s.v_empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s.s_full[i], 1);
                mbar_init(&s.s_empty[i], 128);
                mbar_init(&s.p_full[i], 128);
                mbar_init(&s.p_empty[i], 1);
                mbar_init(&s.o_full[i], 1);
                mbar_init(&s.o_empty[i], 128);
            }
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(&s.tmem_base, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;
    const uint32_t TM_S = 0, TM_O = 256;   // S buffers at columns 0/128, O buffers at 256/320

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int ks = 0, kph = 0, vs = 0, vph = 0, qph = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                const int b = it / T, qt = it - b * T;
                const int64_t mt0 = (int64_t)b * T;
                mbar_wait(&s.q_empty, qph ^ 1);
                mbar_arrive_expect_tx(&s.q_full, 16384u);
                tma_load_1d(s.Q, qkv + ((mt0 + qt) * kblocks + kb_q) * 16384, 16384u, &s.q_full);
                qph ^= 1;
                for (int j = 0; j < T; ++j) {
                    mbar_wait(&s.k_empty[ks], kph ^ 1);
                    mbar_arrive_expect_tx(&s.k_full[ks], 16384u);
                    tma_load_1d(s.K[ks], qkv + ((mt0 + j) * kblocks + kb_q + 1) * 16384, 16384u, &s.k_full[ks]);
                    if (++ks == AT_KV) { ks = 0; kph ^= 1; }
                    mbar_wait(&s.v_empty[vs], vph ^ 1);
                    mbar_arrive_expect_tx(&s.v_full[vs], 16384u);
                    tma_load_1d(s.V[vs], qkv + ((mt0 + j) * kblocks + kb_q + 2) * 16384, 16384u, &s.v_full[vs]);
                    if (++vs == AT_KV) { vs = 0; vph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(128, 128);
            const uint32_t idesc_o = umma_idesc_bf16_bmn(128, 64);
            int ks = 0, kph = 0, vs = 0, vph = 0, qph = 0;
            int g = 0;   // global key-tile counter of this CTA: S/P/O buffers alternate with g, phases with g >> 1
            auto issue_pv = [&](int gg) {
                const int pb = gg & 1, ph = (gg >> 1) & 1;
                mbar_wait(&s.p_full[pb], ph);
                mbar_wait(&s.v_full[vs], vph);
                mbar_wait(&s.o_empty[pb], ph ^ 1);
                tc_fence_after();
                const uint64_t bdesc = umma_desc_sw128_mn(smem_u32(s.V[vs]));
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {   // 128 keys = 8 x 16; A: 2 key-blocks of 64 (16 KiB each), +32 B per step inside
                    const uint64_t adesc = umma_desc_sw128(smem_u32(s.P[pb] + (kk >> 2) * 16384)) + (uint64_t)((kk & 3) * 2);
                    umma_bf16(tmem + TM_O + pb * 64, adesc, bdesc + (uint64_t)(kk * (2048 >> 4)), idesc_o, kk > 0 ? 1u : 0u);
                }
                umma_commit(&s.v_empty[vs]);
                umma_commit(&s.p_empty[pb]);
                umma_commit(&s.o_full[pb]);
                if (++vs == AT_KV) { vs = 0; vph ^= 1; }
            };
            for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
                mbar_wait(&s.q_full, qph);
                for (int j = 0; j < T; ++j, ++g) {
                    const int sb = g & 1, ph = (g >> 1) & 1;
                    mbar_wait(&s.k_full[ks], kph);
                    mbar_wait(&s.s_empty[sb], ph ^ 1);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(s.Q));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(s.K[ks]));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem + TM_S + sb * 128, adesc + kk * 2, bdesc + kk * 2, idesc_s, kk > 0 ? 1u : 0u);
                    umma_commit(&s.k_empty[ks]);
                    umma_commit(&s.s_full[sb]);
                    if (j == T - 1) umma_commit(&s.q_empty);   // Q may be overwritten once the last S of this item is done
                    if (++ks == AT_KV) { ks = 0; kph ^= 1; }
                    if (j > 0) issue_pv(g - 1);
                }
                issue_pv(g - 1);
                qph ^= 1;
            }
        }
        __syncwarp();
    } else {
        // =============================== softmax / epilogue ===============================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        int g = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int b = it / T, qt = it - b * T;
            float m_run = -INFINITY, l_run = 0.0f;
            float o[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) o[c] = 0.0f;
            float alpha_prev = 1.0f;
            for (int j = 0; j < T; ++j, ++g) {
                const int sb = g & 1, ph = (g >> 1) & 1;
                mbar_wait(&s.s_full[sb], ph);
                tc_fence_after();
                // The score tile is read out of TMEM ONCE, into 128 registers (TMEM reads are 64 B/clk per SM: at two passes over
                // S plus the O fold they were 71 us of this kernel's 89), and its buffer goes back to the MMA warp right away.
                uint32_t sr[4][32];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) tmem_ld32(tmem + tlane + TM_S + sb * 128 + q4 * 32, sr[q4]);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) tmem_ld_wait32(sr[q4]);
                tc_fence_before();
                mbar_arrive(&s.s_empty[sb]);
                // row max of the raw scores (four independent chains: this warp is alone on its scheduler, a single 128-long
                // dependent max / add chain would cost 4 cycles per element)
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
#pragma unroll
                    for (int c = 0; c < 32; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(sr[q4][c]));
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                const float m_new = fmaxf(m_run, mx * scale_log2);
                const float alpha = ex2(m_run - m_new);           // 0 on the first tile (m_run = -inf)
                // p = 2^(s*scale - m), written as the bf16 A operand of P.V
                mbar_wait(&s.p_empty[sb], ph ^ 1);
                float sum4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int c0 = q4 * 32;
                    uint8_t* Pt = s.P[sb] + (c0 >> 6) * 16384;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
                        float pv[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            pv[e] = ex2(fmaf(__uint_as_float(sr[q4][qq * 8 + e]), scale_log2, -m_new));
                            sum4[e & 3] += pv[e];
                        }
                        uint4 pk;
                        pk.x = pack_bf16x2(pv[0], pv[1]);
                        pk.y = pack_bf16x2(pv[2], pv[3]);
                        pk.z = pack_bf16x2(pv[4], pv[5]);
                        pk.w = pack_bf16x2(pv[6], pv[7]);
                        *reinterpret_cast<uint4*>(Pt + sw128(row, ((c0 & 63) >> 3) + qq)) = pk;
                    }
                }
                fence_proxy_async();
                mbar_arrive(&s.p_full[sb]);
                // fold O_{j-1} (relative to the previous max) and rescale to the new max
                if (j > 0) {
                    const int ob = (g - 1) & 1, oph = ((g - 1) >> 1) & 1;
                    mbar_wait(&s.o_full[ob], oph);
                    tc_fence_after();
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t r[32];
                        tmem_ld32(tmem + tlane + TM_O + ob * 64 + h * 32, r);
                        tmem_ld_wait32(r);
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[h * 32 + c] = (o[h * 32 + c] + __uint_as_float(r[c])) * alpha;
                    }
                    tc_fence_before();
                    mbar_arrive(&s.o_empty[ob]);
                }
                l_run = l_run * alpha + ((sum4[0] + sum4[1]) + (sum4[2] + sum4[3]));
                m_run = m_new;
                (void)alpha_prev;
            }
            {   // last P.V of this item
                const int ob = (g - 1) & 1, oph = ((g - 1) >> 1) & 1;
                mbar_wait(&s.o_full[ob], oph);
                tc_fence_after();
                const float inv = 1.0f / l_run;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[32];
                    tmem_ld32(tmem + tlane + TM_O + ob * 64 + h * 32, r);
                    tmem_ld_wait32(r);
#pragma unroll
                    for (int c = 0; c < 32; ++c) o[h * 32 + c] = (o[h * 32 + c] + __uint_as_float(r[c])) * inv;
                }
                tc_fence_before();
                mbar_arrive(&s.o_empty[ob]);
            }
            const int64_t mt = (int64_t)b * T + qt;
            if (y_cm) {
                float* dst = y_cm + (int64_t)b * y_bstride + (int64_t)qt * 128 + row;
#pragma unroll
                for (int c = 0; c < 64; ++c) dst[(int64_t)c * N] = o[c];
            }
            if (y_act) {
                uint8_t* t = y_act + (mt * y_kblocks + y_kb) * 16384;
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                    uint4 pk;
                    pk.x = pack_bf16x2(o[qq * 8 + 0], o[qq * 8 + 1]);
                    pk.y = pack_bf16x2(o[qq * 8 + 2], o[qq * 8 + 3]);
                    pk.z = pack_bf16x2(o[qq * 8 + 4], o[qq * 8 + 5]);
                    pk.w = pack_bf16x2(o[qq * 8 + 6], o[qq * 8 + 7]);
                    *reinterpret_cast<uint4*>(t + sw128(row, qq)) = pk;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace gfs

extern "C" int gfs_attention_fwd(const void* qkv_act, int kblocks, int kb_q, int B, int N, float scale, float* y_cm,
                                 int64_t y_bstride, void* y_act, int y_kblocks, int y_kb, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(qkv_act, GFS_ERR_BAD_ARG, "gfs_attention_fwd: null pointer");
    GFS_REQUIRE(y_cm || y_act, GFS_ERR_BAD_ARG, "gfs_attention_fwd: no output requested");
    GFS_REQUIRE(B > 0 && N > 0 && kb_q >= 0 && kb_q + 3 <= kblocks, GFS_ERR_BAD_ARG, "gfs_attention_fwd: bad sizes");
    GFS_REQUIRE(N % 128 == 0, GFS_ERR_UNSUPPORTED, "gfs_attention_fwd: N=%d must be a multiple of 128 (tiles may not straddle blocks)", N);
    GFS_REQUIRE(!y_act || (y_kb >= 0 && y_kb < y_kblocks), GFS_ERR_BAD_ARG, "gfs_attention_fwd: output block out of range");
    GFS_REQUIRE((reinterpret_cast<uintptr_t>(qkv_act) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_act) & 15) == 0, GFS_ERR_BAD_ARG,
                "gfs_attention_fwd: pointers must be 16-byte aligned");
    const int T = N / 128;
    const int items = B * T;
    const int sms = sm_count();
    GFS_REQUIRE(sms > 0, GFS_ERR_CUDA, "gfs_attention_fwd: cannot query the device");
    const size_t smem = sizeof(AtSmem) + 1024;
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(attention_kernel), smem));
    launch_pdl(attention_kernel, dim3((unsigned)(items < sms ? items : sms)), dim3(AT_THREADS), smem, static_cast<cudaStream_t>(stream),
        static_cast<const uint8_t*>(qkv_act), kblocks, kb_q, T, items, scale * 1.4426950408889634f, N, y_cm, y_bstride,
        static_cast<uint8_t*>(y_act), y_kblocks, y_kb);
    GFS_LAUNCH_OK("attention_kernel");
    return GFS_OK;
}
