// rowsel_tc.cu -- the "points x dictionary" contractions of rowsel.cu on tcgen05, with the integer result still bit-exact.
//
//   gfs_gw_project_tc    model/capl.py:344-353   cos = <gp_l2[g], ec>/max(|ec|,1e-12) -> softmax_g(10 cos), argmax_g
//   gfs_kmeans_assign_tc sklearn _k_means_lloyd.pyx:196-218 (behind get_basis.py:210): argmin_c (|c|^2 - 2 x.c)
//
// The (points x D) . (D x <=192) product is a real GEMM (2*192*192 flop per point against 768 bytes): on CUDA cores it
// is FFMA bound (rowsel.cu: 69 % FMA pipe), on the tensor cores it drops under the time HBM needs to deliver the points.
// What must not change is the INTEGER result (GW assignment / k-means label): it is defined by the pinned fp32 fma chain
// of oracle/gfs_oracle.c.  So the tensor cores compute S ~= x.g with a bf16 hi/lo split of both operands
// (x_h.g_h + x_h.g_l + x_l.g_h, fp32 accumulation in TMEM, |S - x.g| <= 2^-13.6 |x||g| even if every rounding aligned),
// every row keeps its best and second best score, and a row whose two best scores are closer than the error bound allows
// is appended to a re-check list: gfs::rowsel_recheck_kernel evaluates the pinned chain for exactly those rows (a warp per
// row) and overwrites their result.  Features (the softmax of the cosines) carry a 2e-2 tolerance and are taken from the
// tensor-core values directly.
//
// Persistent CTA, 16 warps:
//   warps 0-7   producers: fp32 channel-major points (coalesced over the 128 points of the tile) -> bf16 hi / lo operand
//               tiles in shared memory (K-major SWIZZLE_128B, 64 channels per stage), partial squared norms
//   warp  8     TMA: the dictionary's packed hi / lo image (rowsel_pack_kernel), 2 x 24 KiB per stage
//   warp  9     MMA issue (warp-uniform loop, one elected lane): 12 x tcgen05.mma 128x192x16 per stage
//   warps 12-15 epilogue: tcgen05.ld, best / second best, (GW) softmax features as bf16 act tiles, ambiguity test
#include "fp32_tile.cuh"

namespace gfs {

constexpr int RT_ROWS = 128;      // points per tile
constexpr int RT_N = 192;         // dictionary columns of one MMA (zero padded)
constexpr int RT_THREADS = 512;
constexpr uint32_t RT_BTILE = RT_N * 128;   // bytes of one 64-channel dictionary tile (hi or lo)

struct RtSmem {
    uint8_t A[2][2][16384];       // [stage][hi | lo]  128 points x 64 channels
    uint8_t B[2][2][RT_BTILE];    // [stage][hi | lo]  192 entries x 64 channels
    float xn[4][2][RT_ROWS];      // [tile & 3][channel half] partial squared norms
    uint64_t a_full[2], a_empty[2], b_full[2], b_empty[2], accf[2], acce[2];
    uint32_t tmem_base;
};

enum { RT_GW = 0, RT_KMEANS = 1 };

// dict_t (D, Gp) fp32 channel-major -> per 64-channel block: [hi tile | lo tile], 192 rows x 128 B, K-major SWIZZLE_128B.
// Also the largest squared norm of an entry (k-means: scales the ambiguity bound).
__global__ void rowsel_pack_kernel(const float* __restrict__ dict_t, int D, int G, int Gp, uint8_t* __restrict__ img,
                                   float* __restrict__ cmax2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = D >> 3;
    if (i >= RT_N * chunks) return;
    const int j = i / chunks, qq = i - j * chunks;          // entry, 8-channel chunk
    const int kb = qq >> 3, q = qq & 7;
    uint32_t hp[4], lp[4];
    float n2 = 0.0f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = qq * 8 + 2 * e;
        const float v0 = j < G ? dict_t[(int64_t)c * Gp + j] : 0.0f, v1 = j < G ? dict_t[(int64_t)(c + 1) * Gp + j] : 0.0f;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        __nv_bfloat162 hh;
        hh.x = h0;
        hh.y = h1;
        hp[e] = *reinterpret_cast<uint32_t*>(&hh);
        lp[e] = pack_bf16x2(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
        n2 = fmaf(v0, v0, fmaf(v1, v1, n2));
    }
    uint8_t* t = img + (size_t)kb * 2 * RT_BTILE;
    *reinterpret_cast<uint4*>(t + sw128(j, q)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4*>(t + RT_BTILE + sw128(j, q)) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    // an upper bound of max_j |c_j|^2 is all that is needed: sum the chunk maxima's contributions per entry with atomics on
    // the ordered bit pattern of a non-negative float (chunks of one entry add up)
    if (cmax2) atomicAdd(cmax2 + 1 + j, n2);
}
__global__ void rowsel_cmax_kernel(float* __restrict__ cmax2, int G) {      // cmax2[0] = max_j cmax2[1 + j]
    float m = 0.0f;
    for (int j = threadIdx.x; j < G; j += 32) m = fmaxf(m, cmax2[1 + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) cmax2[0] = m;
}

template <int MODE>
__global__ void __launch_bounds__(RT_THREADS, 1)
rowsel_tc_kernel(const float* __restrict__ x, int64_t bstride, int64_t cstride, int D, int N, int64_t M, int ntiles,
                 const uint8_t* __restrict__ img, int G, int Gp, const float* __restrict__ cnorm, const float* __restrict__ cmax2,
                 uint8_t* __restrict__ cos_act, int kblocks, int kb0, float* __restrict__ cos_cm, int32_t* __restrict__ sel,
                 int32_t* __restrict__ recheck, int32_t* __restrict__ recheck_cnt) {
    extern __shared__ unsigned char smem_raw[];
    RtSmem& s = *reinterpret_cast<RtSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = D >> 6;

    if (warp == 9) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s.a_full[i], 8);
                mbar_init(&s.a_empty[i], 1);
                mbar_init(&s.b_full[i], 1);
                mbar_init(&s.b_empty[i], 1);
                mbar_init(&s.accf[i], 1);
                mbar_init(&s.acce[i], 4);
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

    if (warp < 8) {
        // =============================== producers ===============================
        const int r = tid & 127, half = tid >> 7;           // point row, channel half of the 64-channel stage
        int it = 0;                                           // (tile, kb) stage counter
        int lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int64_t m = (int64_t)tile * RT_ROWS + r;
            const bool valid = m < M;
            const float* xp;
            if (MODE == RT_GW) {
                const int64_t b = (valid ? m : 0) / N;
                xp = x + b * bstride + ((valid ? m : 0) - b * N);
            } else {
                xp = x + (valid ? m : 0);
            }
            float nrm = 0.0f;
            float v[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = valid ? __ldg(xp + (int64_t)(half * 32 + c) * cstride) : 0.0f;
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int st = it & 1;
                float vn[32];                                  // the next stage's channels are in flight while this one is converted
                if (kb + 1 < nkb) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) vn[c] = valid ? __ldg(xp + (int64_t)((kb + 1) * 64 + half * 32 + c) * cstride) : 0.0f;
                }
                mbar_wait(&s.a_empty[st], ((it >> 1) & 1) ^ 1);
                uint8_t* ah = s.A[st][0];
                uint8_t* al = s.A[st][1];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t hp[4], lp[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float v0 = v[q * 8 + 2 * e], v1 = v[q * 8 + 2 * e + 1];
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                        __nv_bfloat162 hh;
                        hh.x = h0;
                        hh.y = h1;
                        hp[e] = *reinterpret_cast<uint32_t*>(&hh);
                        lp[e] = pack_bf16x2(v0 - __bfloat162float(h0), v1 - __bfloat162float(h1));
                        nrm = fmaf(v0, v0, fmaf(v1, v1, nrm));
                    }
                    const uint32_t o = sw128(r, half * 4 + q);
                    *reinterpret_cast<uint4*>(ah + o) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                    *reinterpret_cast<uint4*>(al + o) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
                }
                // the partial norm goes out BEFORE the tile's last arrive: that release orders it before the epilogue's read
                if (kb == nkb - 1) s.xn[lt & 3][half][r] = nrm;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s.a_full[st]);
                if (kb + 1 < nkb) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] = vn[c];
                }
            }
        }
    } else if (warp == 8) {
        // =============================== TMA: dictionary tiles ===============================
        if (lane == 0) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int st = it & 1;
                    mbar_wait(&s.b_empty[st], ((it >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&s.b_full[st], 2u * RT_BTILE);
                    tma_load_1d(s.B[st][0], img + (size_t)kb * 2 * RT_BTILE, 2u * RT_BTILE, &s.b_full[st]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // =============================== MMA issue ===============================
        const uint32_t idesc = umma_idesc_bf16(128, RT_N);
        const uint64_t a0 = umma_desc_sw128(smem_u32(s.A[0][0]));
        const uint64_t b0 = umma_desc_sw128(smem_u32(s.B[0][0]));
        constexpr uint32_t A_ST = 2 * 16384 >> 4, A_LO = 16384 >> 4, B_ST = 2 * RT_BTILE >> 4, B_LO = RT_BTILE >> 4;
        int it = 0, lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int acc = lt & 1;
            mbar_wait(&s.acce[acc], ((lt >> 1) & 1) ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int st = it & 1;
                mbar_wait(&s.a_full[st], (it >> 1) & 1);
                mbar_wait(&s.b_full[st], (it >> 1) & 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t aB = a0 + (uint64_t)(st * A_ST), bB = b0 + (uint64_t)(st * B_ST);
                    const uint32_t d = tmem + acc * 256;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        umma_bf16(d, aB + ks * 2, bB + ks * 2, idesc, (kb | ks) ? 1u : 0u);
                        umma_bf16(d, aB + ks * 2, bB + B_LO + ks * 2, idesc, 1u);
                        umma_bf16(d, aB + A_LO + ks * 2, bB + ks * 2, idesc, 1u);
                    }
                    umma_commit(&s.a_empty[st]);
                    umma_commit(&s.b_empty[st]);
                    if (kb == nkb - 1) umma_commit(&s.accf[acc]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 12) {
        // =============================== epilogue: one thread per point ===============================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16);
        const float cmax = MODE == RT_KMEANS ? sqrtf(__ldg(cmax2)) : 0.0f;
        int lt = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++lt) {
            const int acc = lt & 1;
            mbar_wait_backoff(&s.accf[acc], (lt >> 1) & 1, 64);
            tc_fence_after();
            const int64_t m = (int64_t)tile * RT_ROWS + row;
            const bool valid = m < M;
            const float nrm = s.xn[lt & 3][0][row] + s.xn[lt & 3][1][row];
            const uint32_t t0 = tbase + acc * 256;
            uint32_t rr[32];
            float b1, b2;
            int i1 = 0;
            if (MODE == RT_GW) {
                const float inv = 10.0f / fmaxf(sqrtf(nrm), 1e-12f);
                b1 = -INFINITY;
                b2 = -INFINITY;
                for (int ch = 0; ch * 32 < G; ++ch) {
                    tmem_ld32(t0 + ch * 32, rr);
                    tmem_ld_wait32(rr);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float lg = ch * 32 + c < G ? __uint_as_float(rr[c]) * inv : -INFINITY;
                        if (lg > b1) {
                            b2 = b1;
                            b1 = lg;
                            i1 = ch * 32 + c;
                        } else {
                            b2 = fmaxf(b2, lg);
                        }
                    }
                }
                float sum = 0.0f;
                for (int ch = 0; ch * 32 < G; ++ch) {
                    tmem_ld32(t0 + ch * 32, rr);
                    tmem_ld_wait32(rr);
#pragma unroll
                    for (int c = 0; c < 32; ++c) sum += ch * 32 + c < G ? __expf(__uint_as_float(rr[c]) * inv - b1) : 0.0f;
                }
                const float rs = 1.0f / sum;
                for (int ch = 0; ch * 32 < Gp; ++ch) {
                    tmem_ld32(t0 + ch * 32, rr);
                    tmem_ld_wait32(rr);
                    if (valid) {
                        float e[32];
#pragma unroll
                        for (int c = 0; c < 32; ++c) e[c] = ch * 32 + c < G ? __expf(__uint_as_float(rr[c]) * inv - b1) * rs : 0.0f;
                        if (cos_act) {
                            uint8_t* tl = cos_act + (((m >> 7) * kblocks + kb0 + (ch >> 1)) * 16384);
                            const uint32_t rw = (uint32_t)(m & 127);
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                *reinterpret_cast<uint4*>(tl + sw128(rw, (ch & 1) * 4 + q)) =
                                    make_uint4(pack_bf16x2(e[q * 8], e[q * 8 + 1]), pack_bf16x2(e[q * 8 + 2], e[q * 8 + 3]),
                                               pack_bf16x2(e[q * 8 + 4], e[q * 8 + 5]), pack_bf16x2(e[q * 8 + 6], e[q * 8 + 7]));
                        }
                        if (cos_cm) {
                            const int64_t b = m / N, n = m - b * N;
#pragma unroll
                            for (int c = 0; c < 32; ++c)
                                if (ch * 32 + c < G) cos_cm[(b * G + ch * 32 + c) * N + n] = e[c];
                        }
                    }
                }
            } else {
                b1 = INFINITY;
                b2 = INFINITY;
                for (int ch = 0; ch * 32 < G; ++ch) {
                    tmem_ld32(t0 + ch * 32, rr);
                    tmem_ld_wait32(rr);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int col = ch * 32 + c;
                        const float sc = col < G ? fmaf(-2.0f, __uint_as_float(rr[c]), __ldg(cnorm + (col < G ? col : 0))) : INFINITY;
                        if (sc < b1) {
                            b2 = b1;
                            b1 = sc;
                            i1 = col;
                        } else {
                            b2 = fminf(b2, sc);
                        }
                    }
                }
            }
            // hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.acce[acc]);
            if (valid) {
                sel[m] = i1;
                // the two best scores closer than the error bound of the tensor-core product (x 3 safety, see the header):
                // the pinned fp32 chain decides (rowsel_recheck_kernel)
                bool amb;
                if (MODE == RT_GW) amb = !(b1 - b2 > 5e-3f) || !(nrm > 1e-20f);
                else amb = !(b2 - b1 > 0x1p-10f * sqrtf(nrm) * cmax);
                if (amb) recheck[atomicAdd(recheck_cnt, 1)] = (int32_t)m;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// One warp per listed row: the pinned fp32 chains (c ascending) of every dictionary entry, exactly as rowsel.cu evaluates
// them; the winner (ties -> lowest index) overwrites the tensor-core result.
template <int MODE>
__global__ void __launch_bounds__(256)
rowsel_recheck_kernel(const float* __restrict__ x, int64_t bstride, int64_t cstride, int D, int N, const float* __restrict__ dict_t,
                      int G, int Gp, const float* __restrict__ cnorm, const int32_t* __restrict__ recheck,
                      const int32_t* __restrict__ recheck_cnt, int32_t* __restrict__ sel) {
    const int lane = threadIdx.x & 31;
    const int nrows = *recheck_cnt;
    for (int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < nrows; w += gridDim.x * (blockDim.x >> 5)) {
        const int64_t m = recheck[w];
        const float* xp;
        if (MODE == RT_GW) {
            const int64_t b = m / N;
            xp = x + b * bstride + (m - b * N);
        } else {
            xp = x + m;
        }
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float nrm = 0.0f;
        for (int c = 0; c < D; ++c) {
            const float a = __ldg(xp + (int64_t)c * cstride);
            nrm = fmaf(a, a, nrm);
            const float* g = dict_t + (int64_t)c * Gp;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                const int col = j * 32 + lane;
                if (col < G) acc[j] = fmaf(a, __ldg(g + col), acc[j]);
            }
        }
        float best = MODE == RT_GW ? -INFINITY : INFINITY;
        int bi = 0x7fffffff;
        const float inv = 10.0f / fmaxf(sqrtf(nrm), 1e-12f);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const int col = j * 32 + lane;
            if (col < G) {
                const float sc = MODE == RT_GW ? acc[j] * inv : fmaf(-2.0f, acc[j], __ldg(cnorm + col));
                const bool better = MODE == RT_GW ? (sc > best || (sc == best && col < bi)) : (sc < best || (sc == best && col < bi));
                if (better) {
                    best = sc;
                    bi = col;
                }
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const bool better = MODE == RT_GW ? (ob > best || (ob == best && oi < bi)) : (ob < best || (ob == best && oi < bi));
            if (better) {
                best = ob;
                bi = oi;
            }
        }
        if (lane == 0) sel[m] = bi;
    }
}

struct RtPlan {
    size_t off_img, off_cmax, off_cnt, off_list, total;
};
static RtPlan rt_plan(int64_t rows, int D) {
    RtPlan p;
    p.off_img = 0;
    size_t o = (size_t)(D >> 6) * 2 * RT_BTILE;
    p.off_cmax = o;                       // 1 + 192 floats
    o += 256 * 4;
    p.off_cnt = o;
    o += 256;
    p.off_list = o;
    o += (size_t)rows * 4;
    p.total = (o + 255) / 256 * 256;
    return p;
}

template <int MODE>
static int rt_run(const char* who, const float* x, int64_t bstride, int64_t cstride, int D, int N, int64_t M, const float* dict_t, int G,
                  int Gp, const float* cnorm, void* cos_act, int kblocks, int kb0, float* cos_cm, int32_t* sel, void* workspace,
                  int64_t workspace_bytes, cudaStream_t st) {
    const RtPlan p = rt_plan(M, D);
    GFS_REQUIRE(workspace && workspace_bytes >= (int64_t)p.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, GFS_ERR_BAD_ARG,
                "%s: workspace of %lld bytes (256-byte aligned) needed, got %lld", who, (long long)p.total, (long long)workspace_bytes);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* cmax2 = reinterpret_cast<float*>(ws + p.off_cmax);
    int32_t* cnt = reinterpret_cast<int32_t*>(ws + p.off_cnt);
    int32_t* list = reinterpret_cast<int32_t*>(ws + p.off_list);
    GFS_CUDA_OK(cudaMemsetAsync(ws + p.off_cmax, 0, p.off_list - p.off_cmax, st));
    const int pk = RT_N * (D >> 3);
    rowsel_pack_kernel<<<(pk + 255) / 256, 256, 0, st>>>(dict_t, D, G, Gp, ws, cmax2);
    GFS_LAUNCH_OK("rowsel_pack_kernel");
    rowsel_cmax_kernel<<<1, 32, 0, st>>>(cmax2, G);
    GFS_LAUNCH_OK("rowsel_cmax_kernel");
    const int ntiles = (int)((M + RT_ROWS - 1) / RT_ROWS);
    const int sms = sm_count();
    GFS_REQUIRE(sms > 0, GFS_ERR_CUDA, "%s: cannot query the device", who);
    const size_t smem = sizeof(RtSmem) + 1024;
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(rowsel_tc_kernel<MODE>), smem));
    rowsel_tc_kernel<MODE><<<ntiles < sms ? ntiles : sms, RT_THREADS, smem, st>>>(
        x, bstride, cstride, D, N, M, ntiles, ws, G, Gp, cnorm, cmax2, static_cast<uint8_t*>(cos_act), kblocks, kb0, cos_cm, sel, list, cnt);
    GFS_LAUNCH_OK("rowsel_tc_kernel");
    rowsel_recheck_kernel<MODE><<<sms * 2, 256, 0, st>>>(x, bstride, cstride, D, N, dict_t, G, Gp, cnorm, list, cnt, sel);
    GFS_LAUNCH_OK("rowsel_recheck_kernel");
    return GFS_OK;
}

}  // namespace gfs

extern "C" int64_t gfs_rowsel_tc_workspace_bytes(int64_t rows, int D) {
    if (rows <= 0 || D <= 0 || D % 64 != 0 || D > 256) return 0;
    return (int64_t)gfs::rt_plan(rows, D).total;
}

extern "C" int gfs_gw_project_tc(const float* ec, int64_t ec_bstride, int B, int D, int N, const float* gp_l2t, int G, int Gp,
                                 void* cosine_act, int kblocks, int kb0, float* cosine_cm, int32_t* assignment, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(ec && gp_l2t && assignment, GFS_ERR_BAD_ARG, "gfs_gw_project_tc: null pointer");
    GFS_REQUIRE(B > 0 && D > 0 && N > 0 && G > 0, GFS_ERR_BAD_ARG, "gfs_gw_project_tc: non-positive size");
    GFS_REQUIRE(D % 64 == 0 && D <= 256, GFS_ERR_UNSUPPORTED, "gfs_gw_project_tc: D=%d (need a multiple of 64, <= 256; use gfs_gw_project)", D);
    GFS_REQUIRE(N % 128 == 0, GFS_ERR_UNSUPPORTED, "gfs_gw_project_tc: N=%d (need a multiple of 128; use gfs_gw_project)", N);
    GFS_REQUIRE(G <= Gp && Gp <= RT_N && Gp % 64 == 0, GFS_ERR_UNSUPPORTED, "gfs_gw_project_tc: G=%d Gp=%d (need G <= Gp <= 192, Gp %% 64 == 0)", G, Gp);
    GFS_REQUIRE((reinterpret_cast<uintptr_t>(cosine_act) & 15) == 0, GFS_ERR_BAD_ARG, "gfs_gw_project_tc: cosine_act must be 16-byte aligned");
    if (cosine_act) GFS_REQUIRE(kb0 >= 0 && kb0 + Gp / 64 <= kblocks, GFS_ERR_BAD_ARG, "gfs_gw_project_tc: output blocks out of range");
    return rt_run<RT_GW>("gfs_gw_project_tc", ec, ec_bstride, N, D, N, (int64_t)B * N, gp_l2t, G, Gp, nullptr, cosine_act, kblocks, kb0, cosine_cm,
                         assignment, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int gfs_kmeans_assign_tc(const float* xt, int64_t n, int64_t npad, int D, const float* centers_t, int K, int Kp, float* cnorm,
                                    int32_t* labels, void* workspace, int64_t workspace_bytes, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(xt && centers_t && cnorm && labels, GFS_ERR_BAD_ARG, "gfs_kmeans_assign_tc: null pointer");
    GFS_REQUIRE(n > 0 && npad >= n && D > 0 && K > 0, GFS_ERR_BAD_ARG, "gfs_kmeans_assign_tc: non-positive size");
    GFS_REQUIRE(n < (int64_t)1 << 31, GFS_ERR_UNSUPPORTED, "gfs_kmeans_assign_tc: n=%lld exceeds 2^31 per shard", (long long)n);
    GFS_REQUIRE(D % 64 == 0 && D <= 256, GFS_ERR_UNSUPPORTED, "gfs_kmeans_assign_tc: D=%d (need a multiple of 64, <= 256; use gfs_kmeans_assign)", D);
    GFS_REQUIRE(K <= Kp && Kp <= RT_N, GFS_ERR_UNSUPPORTED, "gfs_kmeans_assign_tc: K=%d Kp=%d (need K <= Kp <= 192)", K, Kp);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    sqnorm_kernel<<<dim3((Kp + 255) / 256, 1), 256, 0, st>>>(centers_t, 0, D, Kp, cnorm);
    GFS_LAUNCH_OK("sqnorm_kernel");
    return rt_run<RT_KMEANS>("gfs_kmeans_assign_tc", xt, 0, npad, D, 1, n, centers_t, K, Kp, cnorm, nullptr, 0, 0, nullptr, labels, workspace,
                             workspace_bytes, st);
}
