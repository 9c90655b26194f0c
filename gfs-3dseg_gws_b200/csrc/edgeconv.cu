// edgeconv.cu -- fused EdgeConv given the kNN graph, conv2 on tcgen05 with the accumulator in TMEM.
//
// Replaces model/dgcnn.py:35-41 (gather / x_j - x_i / cat -> (B,2C,N,k)), :53-58 (conv1+BN+LReLU, conv2+BN+LReLU
// over (B,64,N,k)) and :118 (max over k).  With the first conv split per point (pointwise.cu) an edge's hidden
// vector is h1 = LReLU(P'[j] + Q'[i]); a tile is 128 points x ONE neighbour slot, so the max over the k slots is an
// element-wise max across k accumulator tiles held by the same thread -- no cross-lane reduction, and because BN2's
// scale is folded into W2 the remaining "+shift, LeakyReLU" is monotone and is applied once after the max.
//
// Warp roles (persistent CTA, one per SM):
//   warps 0-7  producers : cp.async gather of the bf16 row P'[idx] - mu (128 B contiguous per edge, gfs_edge_pq_f32) into a
//                          shared-memory ring, EC_NG - 1 slots ahead; + (Q' + mu) (fp32, registers) -> LReLU -> bf16 ->
//                          SWIZZLE_128B A tile in shared memory ->
//                          fence.proxy.async -> arrive full[stage]
//   warp  12   MMA       : one thread issues 4 x tcgen05.mma (128x64x16) per A tile into TMEM stage acc;
//                          tcgen05.commit -> empty[stage], accf[acc].  W2 (8 KB image) arrives once by TMA bulk copy.
//   warps 8-11 epilogue  : tcgen05.ld 32x32b, running max in registers, arrive acce[acc]; after slot k-1:
//                          +shift, LeakyReLU, store fp32 channel-major and bf16 "act" tiles.
#include "common.cuh"

namespace gfs {

constexpr int EC_TM = 128;
constexpr int EC_NST = 4;
constexpr int EC_NACC = 4;
constexpr int EC_PROD = 512;        // producer threads (warps 0-15): the producers are latency bound, 4 per scheduler hide it
constexpr int EC_PW = EC_PROD / 32;
constexpr int EC_THREADS = EC_PROD + 128 + 128;  // + 4 epilogue warps (16-19) + the MMA warp (20) and 3 idle warps (one warpgroup: setmaxnreg)
constexpr uint32_t EC_TMEM_COLS = 256;

constexpr int EC_NG = 8;           // gather ring depth (slots of 128 rows x 128 B of bf16 P'); EC_NG - 1 slots are in flight

struct EcSmem {
    uint8_t A[EC_NST][16384];
    uint8_t G[EC_NG][16384];
    uint8_t W[8192];
    float shift[64];
    uint64_t full[EC_NST], empty[EC_NST], accf[EC_NACC], acce[EC_NACC], wbar;
    uint32_t tmem_base;
};

template <bool ARGMAX>
__global__ void __launch_bounds__(EC_THREADS, 1)   // 80 registers at launch; setmaxnreg: producers 64, epilogue 128, MMA group 56
edgeconv_kernel(const uint8_t* __restrict__ pb, const float* __restrict__ qv, const int32_t* __restrict__ idx, const uint8_t* __restrict__ w2p,
                const float* __restrict__ shift2, int N, int k, int64_t M, int ntiles, float* __restrict__ y_cm,
                int64_t y_bstride, uint8_t* __restrict__ y_act, int act_kblocks, int act_kb, uint8_t* __restrict__ y_act2,
                int act2_kblocks, int act2_kb, uint8_t* __restrict__ argmax) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    // align inside the shared window with pointer arithmetic on smem_raw (keeps the .shared address space: LDS/STS, not generic LD/ST)
    EcSmem& s = *reinterpret_cast<EcSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < 64) s.shift[tid] = shift2[tid];
    if (warp == EC_PW + 4) {
        if (lane == 0) {
            for (int i = 0; i < EC_NST; ++i) {
                mbar_init(&s.full[i], EC_PW);        // one arrive per producer WARP: hundreds of arrives on one mbarrier serialise
                mbar_init(&s.empty[i], 1);
            }
            for (int i = 0; i < EC_NACC; ++i) {
                mbar_init(&s.accf[i], 1);
                mbar_init(&s.acce[i], 4);
            }
            mbar_init(&s.wbar, 1);
            mbar_fence_init();
            mbar_arrive_expect_tx(&s.wbar, 8192);
            tma_load_1d(s.W, w2p, 8192, &s.wbar);
        }
        __syncwarp();
        tmem_alloc(&s.tmem_base, EC_TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;

    if (warp < EC_PW) {
        // =============================== producers ===============================
        reg_dealloc<64>();
        // Gathers are cp.async (LDGSTS) copies into an EC_NG-deep shared-memory ring, EC_NG - 1 neighbour slots ahead of the
        // slot being converted, so ~100 KB of P' rows are in flight per SM without holding them in registers.  Each thread
        // converts exactly the 16-byte pieces it copied itself, so cp.async.wait_group is the only synchronisation.
        // The loop is issue bound (ncu: the epilogue warps wait for the producers), so it is written for instruction count:
        // 32-bit row offsets (one IMAD.WIDE per copy), rows past M clamped instead of predicated, packed fp32 add / mul.
        const int q = tid & 7, rsub = tid >> 3;   // rsub 0..63: rows rsub, rsub+64
        int* idxs = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(&s) + sizeof(EcSmem));   // [128][k]
        int stage = 0, phase = 0;
        const uint32_t gq = sw128(rsub, q);                     // the thread's piece inside a slot: rows rsub + 64 p share r & 7
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t m0 = (int64_t)tile * EC_TM;
            named_bar_sync(2, EC_PROD);                          // previous tile's idx no longer needed
            {
                const int64_t lim = (M - m0) * k;                // valid ints of this tile
                const int32_t* src = idx + m0 * k;
                for (int i = tid; i < EC_TM * k; i += EC_PROD) idxs[i] = i < lim ? __ldg(src + i) : 0;
            }
            float2 Q[2][4];
            uint32_t base[2];   // first row of the point's block (rows past M: the last valid point, its result is not stored)
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                int64_t m = m0 + p * 64 + rsub;
                m = m < M ? m : M - 1;
                base[p] = (uint32_t)((m / N) * N);
                const float4* src = reinterpret_cast<const float4*>(qv + m * 64 + q * 8);
                const float4 q0 = __ldg(src), q1 = __ldg(src + 1);
                Q[p][0] = make_float2(q0.x, q0.y);
                Q[p][1] = make_float2(q0.z, q0.w);
                Q[p][2] = make_float2(q1.x, q1.y);
                Q[p][3] = make_float2(q1.z, q1.w);
            }
            named_bar_sync(2, EC_PROD);                          // idx tile visible to all producer threads

            auto issue = [&](int kk) {
                if (kk < k) {
                    unsigned char* G = s.G[kk % EC_NG] + gq;
                    int j[2];
#pragma unroll
                    for (int p = 0; p < 2; ++p) j[p] = idxs[(p * 64 + rsub) * k + kk];
#pragma unroll
                    for (int p = 0; p < 2; ++p)
                        cp_async16(G + p * 8192, pb + (size_t)((base[p] + (uint32_t)j[p]) * 128u + (uint32_t)q * 16u), 16);
                }
                cp_async_commit();                               // always commit: keeps the group count uniform
            };
#pragma unroll
            for (int kk = 0; kk < EC_NG - 1; ++kk) issue(kk);
            for (int kk = 0; kk < k; ++kk) {
                issue(kk + EC_NG - 1);
                cp_async_wait<EC_NG - 1>();                      // this thread's copies of slot kk have landed
                mbar_wait(&s.empty[stage], phase ^ 1);
                uint8_t* A = s.A[stage] + gq;
                const unsigned char* G = s.G[kk % EC_NG] + gq;
                uint4 g4[2];
#pragma unroll
                for (int p = 0; p < 2; ++p) g4[p] = *reinterpret_cast<const uint4*>(G + p * 8192);     // both loads in flight: 8 bf16 channels of P'[j] - mu each
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    const uint32_t gw[4] = {g4[p].x, g4[p].y, g4[p].z, g4[p].w};
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // bf16 -> fp32 is a 16-bit shift (the low half of a word is the even channel); packed add and multiply
                        const float2 v = __fadd2_rn(make_float2(__uint_as_float(gw[e] << 16), __uint_as_float(gw[e] & 0xffff0000u)), Q[p][e]);
                        const float2 t = __fmul2_rn(v, make_float2(0.2f, 0.2f));
                        o[e] = pack_bf16x2(fmaxf(v.x, t.x), fmaxf(v.y, t.y));
                    }
                    *reinterpret_cast<uint4*>(A + p * 8192) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.full[stage]);
                if (++stage == EC_NST) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            cp_async_wait<0>();
        }
    } else if (warp >= EC_PW + 4) {
        // =============================== MMA issuer (warp 20; warps 21-23 only complete its warpgroup) ===============================
        reg_dealloc<56>();
        if (warp == EC_PW + 4 && lane == 0) {
            mbar_wait(&s.wbar, 0);
            const uint32_t idesc = umma_idesc_bf16(128, 64);
            const uint64_t bdesc = umma_desc_sw128(smem_u32(s.W));
            int stage = 0, phase = 0, acc = 0, aphase = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kk = 0; kk < k; ++kk) {
                    mbar_wait(&s.acce[acc], aphase ^ 1);
                    mbar_wait(&s.full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(s.A[stage]));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)   // K = 64 = 4 x 16; +32 bytes per step inside the 128 B swizzle row
                        umma_bf16(tmem + acc * 64, adesc + ks * 2, bdesc + ks * 2, idesc, ks > 0 ? 1u : 0u);
                    umma_commit(&s.empty[stage]);
                    umma_commit(&s.accf[acc]);
                    if (++stage == EC_NST) {
                        stage = 0;
                        phase ^= 1;
                    }
                    if (++acc == EC_NACC) {
                        acc = 0;
                        aphase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ===============================
        reg_alloc<128>();
        const int quarter = warp - EC_PW;        // == warp % 4: the TMEM lane quarter this warp may read
        const int row = quarter * 32 + lane;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        int acc = 0, aphase = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            float mx[64];
            uint32_t am[ARGMAX ? 16 : 1];
#pragma unroll
            for (int c = 0; c < 64; ++c) mx[c] = -INFINITY;
            if (ARGMAX) {
#pragma unroll
                for (int c = 0; c < 16; ++c) am[c] = 0u;
            }
            for (int kk = 0; kk < k; ++kk) {
                mbar_wait_backoff(&s.accf[acc], aphase, 32);     // the epilogue is not the critical path: do not poll at full rate
                tc_fence_after();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[32];
                    tmem_ld32(tmem + tlane + acc * 64 + h * 32, r);
                    tmem_ld_wait32(r);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float vv = __uint_as_float(r[c]);
                        if (ARGMAX) {
                            if (vv > mx[h * 32 + c]) {
                                mx[h * 32 + c] = vv;
                                const int cc = h * 32 + c;
                                am[cc >> 2] = (am[cc >> 2] & ~(0xffu << ((cc & 3) * 8))) | ((uint32_t)kk << ((cc & 3) * 8));
                            }
                        } else {
                            mx[h * 32 + c] = fmaxf(mx[h * 32 + c], vv);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.acce[acc]);
                if (++acc == EC_NACC) {
                    acc = 0;
                    aphase ^= 1;
                }
            }
            const int64_t m = (int64_t)tile * EC_TM + row;
            if (m < M) {
#pragma unroll
                for (int c = 0; c < 64; ++c) mx[c] = lrelu02(mx[c] + s.shift[c]);
                if (y_cm) {
                    const int64_t b = m / N, n = m - b * N;
                    float* o = y_cm + b * y_bstride + n;
#pragma unroll
                    for (int c = 0; c < 64; ++c) o[(int64_t)c * N] = mx[c];
                }
                uint4 pk[8];
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                    pk[qq].x = pack_bf16x2(mx[qq * 8 + 0], mx[qq * 8 + 1]);
                    pk[qq].y = pack_bf16x2(mx[qq * 8 + 2], mx[qq * 8 + 3]);
                    pk[qq].z = pack_bf16x2(mx[qq * 8 + 4], mx[qq * 8 + 5]);
                    pk[qq].w = pack_bf16x2(mx[qq * 8 + 6], mx[qq * 8 + 7]);
                }
                if (y_act) {
                    uint8_t* t = y_act + ((int64_t)tile * act_kblocks + act_kb) * 16384;
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) *reinterpret_cast<uint4*>(t + sw128(row, qq)) = pk[qq];
                }
                if (y_act2) {
                    uint8_t* t = y_act2 + ((int64_t)tile * act2_kblocks + act2_kb) * 16384;
#pragma unroll
                    for (int qq = 0; qq < 8; ++qq) *reinterpret_cast<uint4*>(t + sw128(row, qq)) = pk[qq];
                }
                if (ARGMAX) {
                    uint4* a = reinterpret_cast<uint4*>(argmax + m * 64);
#pragma unroll
                    for (int c = 0; c < 4; ++c) a[c] = make_uint4(am[c * 4], am[c * 4 + 1], am[c * 4 + 2], am[c * 4 + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == EC_PW + 4) {
        tc_fence_after();
        tmem_dealloc(tmem, EC_TMEM_COLS);
    }
}

}  // namespace gfs

extern "C" int gfs_edgeconv_fwd(const void* pb, const float* q, const int32_t* idx, const void* w2_packed, const float* shift2, int B, int N,
                                int k, float* y_cm, int64_t y_bstride, void* y_act, int y_act_kblocks, int y_act_kb,
                                void* y_act2, int y_act2_kblocks, int y_act2_kb, uint8_t* argmax, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(pb && q && idx && w2_packed && shift2, GFS_ERR_BAD_ARG, "gfs_edgeconv_fwd: null pointer");
    GFS_REQUIRE(B > 0 && N > 0 && k > 0, GFS_ERR_BAD_ARG, "gfs_edgeconv_fwd: non-positive size");
    GFS_REQUIRE(k <= 64, GFS_ERR_UNSUPPORTED, "gfs_edgeconv_fwd: k=%d > 64 is not built", k);
    GFS_REQUIRE((int64_t)B * N < ((int64_t)1 << 25), GFS_ERR_UNSUPPORTED, "gfs_edgeconv_fwd: B*N=%lld points exceed the 32-bit gather offsets",
                (long long)B * N);
    GFS_REQUIRE(y_cm || y_act || y_act2, GFS_ERR_BAD_ARG, "gfs_edgeconv_fwd: no output requested");
    GFS_REQUIRE((reinterpret_cast<uintptr_t>(pb) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(w2_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y_act) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_act2) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(argmax) & 15) == 0,
                GFS_ERR_BAD_ARG, "gfs_edgeconv_fwd: pointers must be 16-byte aligned");
    const int64_t M = (int64_t)B * N;
    const int ntiles = (int)((M + EC_TM - 1) / EC_TM);
    const int sms = sm_count();
    GFS_REQUIRE(sms > 0, GFS_ERR_CUDA, "gfs_edgeconv_fwd: cannot query the device");
    const int grid = ntiles < sms ? ntiles : sms;
    const size_t smem = sizeof(EcSmem) + 1024 + (size_t)EC_TM * k * sizeof(int);
    constexpr size_t EC_MAX_SMEM = 227 * 1024;
    GFS_REQUIRE(smem <= EC_MAX_SMEM, GFS_ERR_UNSUPPORTED, "gfs_edgeconv_fwd: k=%d needs %zu bytes of shared memory (> 227 KB)", k, smem);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (argmax) {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(edgeconv_kernel<true>), EC_MAX_SMEM));
        launch_pdl(edgeconv_kernel<true>, grid, dim3(EC_THREADS), smem, st,
        static_cast<const uint8_t*>(pb), q, idx, static_cast<const uint8_t*>(w2_packed), shift2, N, k, M, ntiles, y_cm, y_bstride,
            static_cast<uint8_t*>(y_act), y_act_kblocks, y_act_kb, static_cast<uint8_t*>(y_act2), y_act2_kblocks, y_act2_kb,
            argmax);
    } else {
        GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(edgeconv_kernel<false>), EC_MAX_SMEM));
        launch_pdl(edgeconv_kernel<false>, grid, dim3(EC_THREADS), smem, st,
        static_cast<const uint8_t*>(pb), q, idx, static_cast<const uint8_t*>(w2_packed), shift2, N, k, M, ntiles, y_cm, y_bstride,
            static_cast<uint8_t*>(y_act), y_act_kblocks, y_act_kb, static_cast<uint8_t*>(y_act2), y_act2_kblocks, y_act2_kb,
            nullptr);
    }
    GFS_LAUNCH_OK("edgeconv_kernel");
    return GFS_OK;
}
