// linear.cu -- Conv1d(kernel 1) + BatchNorm(eval, folded) + activation as a tcgen05 bf16 GEMM with a fused epilogue.
//
// Replaces the cuDNN 1x1-conv + ATen batch-norm + elementwise chains of model/dgcnn.py:63-80,121-122 (MLP 192->512->256),
// model/capl.py:435-457 (BaseLearner), model/attention.py:25-27 (q/k/v maps) and model/capl.py:63-65 (fusion).
//
//   Y[m, n] = act( sum_k X[m, k] * Wp[n, k] + shift[n] )        X: bf16 "act" tiles, Wp: packed bf16 (BN scale folded)
//
// Both operands are stored in memory already in the UMMA K-major SWIZZLE_128B arrangement (include/gfs3d.h), so a
// pipeline stage is two plain TMA bulk copies (cp.async.bulk, 16 KiB of X and NT*128 B of W) that land ready to be
// consumed by tcgen05.mma -- no tensor maps, no software swizzle on the load side.
//
// Persistent CTA, 10 warps:  warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2-5 and 6-9 = two epilogue
// groups.  TMEM holds two accumulator stages of NT <= 256 fp32 columns; group g drains stage g (the CTA's even / odd work
// items), so two items are in their epilogue at once (two warps per scheduler hide the TMEM-read and store latencies of
// each other: one group left the issue slots 86 % idle) while the MMAs of the next item run.
#include "common.cuh"

namespace gfs {

constexpr int LN_NST = 4;
constexpr int LN_THREADS = 320;

struct LnSmem {
    uint8_t X[LN_NST][16384];
    uint8_t W[LN_NST][32768];
    float shift[2][256];
    uint64_t full[LN_NST], empty[LN_NST], accf[2], acce[2], wfull;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(LN_THREADS, 1)
linear_kernel(const uint8_t* __restrict__ x_act, int x_kblocks, int x_kb0, int kb_count, const uint8_t* __restrict__ wp,
              const float* __restrict__ shift, int Nout, int NT, int act, int N, int64_t M, int n_mtiles,
              uint8_t* __restrict__ y_act, int y_kblocks, int y_kb0, float* __restrict__ y_cm, int64_t y_bstride, int resident) {
    pdl_enter();
    extern __shared__ unsigned char smem_raw[];
    // align inside the shared window with pointer arithmetic on smem_raw (keeps the .shared address space: LDS/STS, not generic LD/ST)
    LnSmem& s = *reinterpret_cast<LnSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_ntiles = Nout / NT;
    const int n_items = n_mtiles * n_ntiles;
    const uint32_t w_bytes = (uint32_t)NT * 128u;
    // resident mode (the layer's weight slice of one N tile fits the W area: kb_count * NT * 128 B <= 128 KiB): a CTA keeps ONE
    // N tile's weights in shared memory for its whole life and streams only X tiles; its items are (mt, nt fixed).  The
    // streaming mode re-reads the weights with every M tile -- 100 of the 147 MB of L2 traffic of the 192 -> 512 layer.
    const int nt_res = blockIdx.x % n_ntiles, it_first = resident ? blockIdx.x / n_ntiles : blockIdx.x;
    const int it_step = resident ? gridDim.x / n_ntiles : gridDim.x, it_end = resident ? n_mtiles : n_items;
    uint8_t* const Wres = s.W[0];

    if (warp == 1) {
        if (lane == 0) {
            for (int i = 0; i < LN_NST; ++i) {
                mbar_init(&s.full[i], 1);
                mbar_init(&s.empty[i], 1);
            }
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s.accf[i], 1);
                mbar_init(&s.acce[i], 128);
            }
            mbar_init(&s.wfull, 1);
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(&s.tmem_base, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s.tmem_base;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (lane == 0) {
            int stage = 0, phase = 0;
            if (resident && it_first < it_end) {
                mbar_arrive_expect_tx(&s.wfull, w_bytes * (uint32_t)kb_count);
                for (int kb = 0; kb < kb_count; ++kb)
                    tma_load_1d(Wres + (size_t)kb * w_bytes, wp + ((int64_t)kb * Nout + (int64_t)nt_res * NT) * 128, w_bytes, &s.wfull);
            }
            for (int it = it_first; it < it_end; it += it_step) {
                const int mt = resident ? it : it / n_ntiles, nt = resident ? nt_res : it - mt * n_ntiles;
                for (int kb = 0; kb < kb_count; ++kb) {
                    mbar_wait(&s.empty[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&s.full[stage], resident ? 16384u : 16384u + w_bytes);
                    tma_load_1d(s.X[stage], x_act + ((int64_t)mt * x_kblocks + x_kb0 + kb) * 16384, 16384u, &s.full[stage]);
                    if (!resident)
                        tma_load_1d(s.W[stage], wp + ((int64_t)kb * Nout + (int64_t)nt * NT) * 128, w_bytes, &s.full[stage]);
                    if (++stage == LN_NST) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)NT);
            int stage = 0, phase = 0, acc = 0, aphase = 0;
            if (resident && it_first < it_end) mbar_wait(&s.wfull, 0);
            for (int it = it_first; it < it_end; it += it_step) {
                mbar_wait(&s.acce[acc], aphase ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < kb_count; ++kb) {
                    mbar_wait(&s.full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(s.X[stage]));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(resident ? Wres + (size_t)kb * w_bytes : s.W[stage]));
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(tmem + acc * 256, adesc + ks * 2, bdesc + ks * 2, idesc, (kb | ks) ? 1u : 0u);
                    umma_commit(&s.empty[stage]);
                    if (++stage == LN_NST) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&s.accf[acc]);
                if (++acc == 2) {
                    acc = 0;
                    aphase ^= 1;
                }
            }
        }
        __syncwarp();
    } else {
        // =============================== epilogue ===============================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int et = (tid - 64) & 127;   // 0..127 inside the group
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        const int acc = (warp - 2) >> 2;   // the group's accumulator stage
        int aphase = 0;
        if (resident) {                                      // the N tile is fixed: its shifts are loaded once, not once per item
            for (int c = et; c < NT; c += 128) s.shift[acc][c] = shift ? shift[nt_res * NT + c] : 0.0f;
            named_bar_sync(1 + acc, 128);
        }
        for (int it = it_first + acc * it_step; it < it_end; it += 2 * it_step) {
            const int mt = resident ? it : it / n_ntiles, nt = resident ? nt_res : it - mt * n_ntiles;
            if (!resident) {
                named_bar_sync(1 + acc, 128);          // the group's previous item has been read out before its shifts are replaced
                for (int c = et; c < NT; c += 128) s.shift[acc][c] = shift ? shift[nt * NT + c] : 0.0f;
                named_bar_sync(1 + acc, 128);
            }
            mbar_wait(&s.accf[acc], aphase);
            tc_fence_after();
            const int64_t m = (int64_t)mt * 128 + row;
            const int64_t b = m / N, n = m - b * N;
            // one 32-column chunk of the accumulator: + shift, activation, fp32 channel-major and / or bf16 act-tile stores
            auto emit = [&](int c0, const uint32_t (&r)[32]) {
                float v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    float t = __uint_as_float(r[c]) + s.shift[acc][c0 + c];
                    if (act == GFS_ACT_LRELU02) t = lrelu02(t);
                    else if (act == GFS_ACT_RELU) t = fmaxf(t, 0.0f);
                    v[c] = t;
                }
                if (m < M) {
                    if (y_cm) {
                        float* o = y_cm + b * y_bstride + (int64_t)(nt * NT + c0) * N + n;
#pragma unroll
                        for (int c = 0; c < 32; ++c) o[(int64_t)c * N] = v[c];
                    }
                    if (y_act) {
                        const int col = nt * NT + c0;   // column inside this layer's output
                        uint8_t* t = y_act + ((int64_t)mt * y_kblocks + y_kb0 + (col >> 6)) * 16384;
                        const int q0 = (col & 63) >> 3;
#pragma unroll
                        for (int qq = 0; qq < 4; ++qq) {
                            uint4 pk;
                            pk.x = pack_bf16x2(v[qq * 8 + 0], v[qq * 8 + 1]);
                            pk.y = pack_bf16x2(v[qq * 8 + 2], v[qq * 8 + 3]);
                            pk.z = pack_bf16x2(v[qq * 8 + 4], v[qq * 8 + 5]);
                            pk.w = pack_bf16x2(v[qq * 8 + 6], v[qq * 8 + 7]);
                            *reinterpret_cast<uint4*>(t + sw128(row, q0 + qq)) = pk;
                        }
                    }
                }
            };
            // the TMEM read of chunk i+1 is in flight while chunk i is converted and stored (two register buffers taking turns:
            // the ncu source page showed the first use after every tcgen05.wait::ld as the epilogue's largest stall)
            const uint32_t tacc = tmem + tlane + acc * 256;
            uint32_t ra[32], rb[32];
            tmem_ld32(tacc, ra);
            for (int c0 = 0; c0 < NT; c0 += 64) {
                tmem_ld_wait32(ra);
                if (c0 + 32 < NT) tmem_ld32(tacc + c0 + 32, rb);
                emit(c0, ra);
                if (c0 + 32 < NT) {
                    tmem_ld_wait32(rb);
                    if (c0 + 64 < NT) tmem_ld32(tacc + c0 + 64, ra);
                    emit(c0 + 32, rb);
                }
            }
            tc_fence_before();
            mbar_arrive(&s.acce[acc]);
            aphase ^= 1;
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

extern "C" int gfs_linear_bf16(const void* x_act, int x_kblocks, int x_kb0, int kb_count, const void* w_packed,
                               const float* shift, int Nout, int act, int B, int N, void* y_act, int y_kblocks, int y_kb0,
                               float* y_cm, int64_t y_bstride, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x_act && w_packed, GFS_ERR_BAD_ARG, "gfs_linear_bf16: null pointer");
    GFS_REQUIRE(y_act || y_cm, GFS_ERR_BAD_ARG, "gfs_linear_bf16: no output requested");
    GFS_REQUIRE(B > 0 && N > 0 && kb_count > 0 && x_kb0 >= 0 && x_kb0 + kb_count <= x_kblocks, GFS_ERR_BAD_ARG,
                "gfs_linear_bf16: bad sizes (kb0=%d count=%d kblocks=%d)", x_kb0, kb_count, x_kblocks);
    GFS_REQUIRE(Nout > 0 && Nout % 32 == 0, GFS_ERR_UNSUPPORTED, "gfs_linear_bf16: Nout=%d must be a multiple of 32", Nout);
    GFS_REQUIRE(act >= GFS_ACT_NONE && act <= GFS_ACT_RELU, GFS_ERR_BAD_ARG, "gfs_linear_bf16: unknown activation %d", act);
    int NT = Nout <= 256 ? Nout : 256;
    while (Nout % NT) NT -= 32;
    GFS_REQUIRE(NT >= 32 && NT % 16 == 0, GFS_ERR_UNSUPPORTED, "gfs_linear_bf16: cannot tile Nout=%d", Nout);
    if (y_act) {
        GFS_REQUIRE(Nout % 64 == 0 || Nout < 64, GFS_ERR_UNSUPPORTED, "gfs_linear_bf16: bf16 output needs Nout %% 64 == 0");
        GFS_REQUIRE(y_kb0 >= 0 && y_kb0 + (Nout + 63) / 64 <= y_kblocks, GFS_ERR_BAD_ARG, "gfs_linear_bf16: output blocks out of range");
    }
    GFS_REQUIRE((reinterpret_cast<uintptr_t>(x_act) & 15) == 0 && (reinterpret_cast<uintptr_t>(w_packed) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(y_act) & 15) == 0,
                GFS_ERR_BAD_ARG, "gfs_linear_bf16: pointers must be 16-byte aligned");
    const int64_t M = (int64_t)B * N;
    const int n_mtiles = (int)((M + 127) / 128);
    const int items = n_mtiles * (Nout / NT);
    const int sms = sm_count();
    GFS_REQUIRE(sms > 0, GFS_ERR_CUDA, "gfs_linear_bf16: cannot query the device");
    const size_t smem = sizeof(LnSmem) + 1024;
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(linear_kernel), smem));
    const int n_ntiles = Nout / NT;
    int grid = items < sms ? items : sms;
    // resident weights: the N tile's slice must fit the 128 KiB W area; the grid is a multiple of the N tiles so that a CTA's tile is fixed
    const int resident = ((size_t)kb_count * NT * 128 <= sizeof(LnSmem::W)) && grid >= n_ntiles;
    if (resident) grid -= grid % n_ntiles;
    launch_pdl(linear_kernel, grid, dim3(LN_THREADS), smem, static_cast<cudaStream_t>(stream),
        static_cast<const uint8_t*>(x_act), x_kblocks, x_kb0, kb_count, static_cast<const uint8_t*>(w_packed), shift, Nout, NT,
        act, N, M, n_mtiles, static_cast<uint8_t*>(y_act), y_kblocks, y_kb0, y_cm, y_bstride, resident);
    GFS_LAUNCH_OK("linear_kernel");
    return GFS_OK;
}
