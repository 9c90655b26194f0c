// knn_tc.cu -- kNN graph with a tensor-core candidate filter and an exact fp32 finish (model/dgcnn.py:17-23).
//
// The pinned result (oracle/gfs_oracle.c: d = fmaf(2, dot, -xx_i) - xx_j with dot one fp32 fma chain, the k largest per row,
// ties -> ascending index) cannot be formed on tensor cores, but the tensor cores can tell cheaply which few candidates
// can possibly be among the k best:
//
//   1. knn_prep_kernel: distances do not change under a translation, so the filter works on coordinates shifted by a
//      point inside the block's cloud (the mean of four spread-out points; small norms = small absolute error).  Every
//      shifted coordinate is split into two bf16 terms x~ = h + l (+ r, |r| <= 2^-18 |x~|) and stored as [h | l] in the
//      UMMA K-major SWIZZLE_128B tile layout, next to the per-point filter constants
//          nh_j = -|x~_j|^2/2 + a_j (rounded up),  nl_j = -|x~_j|^2/2 - a_j (rounded down),  tag_j = (j << 16) | bf16(2 a_j)
//      and a point-major fp32 copy of the ORIGINAL coordinates.
//   2. knn_tc_kernel: one CTA owns 256 query rows of one block (two UMMA M=128 tiles, operands resident in shared
//      memory) and streams the block's candidates TWICE in stages of 64 (TMA bulk copies, 4 stages in flight).  Per stage
//      three tcgen05.mma chains  h_i.h_j + h_i.l_j + l_i.h_j  leave  D' ~= x~_i.x~_j  in TMEM (fp32, four stages).
//      Sixteen selection warps read TMEM with tcgen05.ld, TWO THREADS PER QUERY ROW (one per 32-column half of a stage),
//      no shuffles, no sorted list, no data-dependent control flow in the scan:
//        pass A  every thread folds the LOWER bounds  w = D' + nl_j  of its candidates into NG running group maxima
//                (group = column position; one FADD + one FMNMX per candidate).  The k-th largest of the row's 2 NG group
//                maxima is a valid lower bound T of the k-th largest w of the row: k different candidates reach it.
//                (sorted in registers by a bitonic network, the two halves meet through shared memory.)
//        pass B  the candidates whose UPPER bound  u = D' + nh_j  reaches  T - 2 a_i  are appended to the row's survivor
//                records {u, tag} in global memory with predicated stores: about 1.2 % of the candidates at NG = 32
//                (N = 2048, k = 20: ~25 per row), independent of the data's order.
//      Error bound (per PAIR, so that one far-away point does not loosen the filter for every row):
//        |v(i,j) - D(i,j)| <= a_i + a_j,   a = 2^-15 |x~|^2 + 2^-25 A + 6 2^-25 |x|^2,  A = sum_c (C - c) x_c^2   per point
//      (bf16 split residual <= 3 * 2^-18 |x~_i||x~_j|, <= 192 fp32 accumulations in the tensor core, one fp32 addition,
//      |x~_i||x~_j| <= (|x~_i|^2+|x~_j|^2)/2, plus the rounding of the pinned fp32 chain on the original coordinates.
//      The chain rounds every partial sum s_c = fl(s_{c-1} + x_ic x_jc) once, |error_c| <= 2^-24 sum_{c'<=c} |x_ic' x_jc'|:
//      channel c' enters C - c' partial sums, so the dot product is off by at most 2^-24 sum_c (C - c) |x_ic x_jc|
//      <= 2^-25 (A_i + A_j): on the filter's scale v ~ D / 2 that is 2^-25 A per point.  The square norms the pinned
//      formula subtracts are COMPUTED numbers, known per point when the operands are prepared: the filter constant of
//      candidate j carries dj = (|x_j|^2 - xx_j) / 2 (|x_j|^2 from an fp64 chain), so their rounding is not an uncertainty
//      at all.  The fmaf / subtraction that finish d, the (1 + 2^-24)^C growth of the partial sums and the 2^-53 of the
//      fp64 chain are inside 6 2^-25 |x|^2.  A <= C |x|^2, about half of it for evenly spread channels: the plain worst
//      case "C roundings of the full product, twice" ((2C+6) 2^-25 |x|^2) is four times as wide.
//      tests/test_gpu_knn_tc.py measures the observed error against the bound).  u is an UPPER bound of D up to the row constant
//      a_i, w a LOWER bound.  The k-th largest lower bound minus 2 a_i cannot exceed the exact k-th best D, hence a
//      candidate with u below T - 2 a_i cannot be among the exact top k.
//   3. knn_finish_kernel: the survivors get their EXACT pinned distance from the point-major fp32 copy, one warp per row
//      and one lane per survivor (32 per round), are ranked by (d desc, index asc) and written out: bit-identical to knn.cu.
//      In SET mode (what the fused EdgeConv needs: a max over the neighbours does not care about their order) the
//      survivors are first classified by their bounds -- certainly in, certainly out, undecided band -- and only the band
//      gets the exact arithmetic; the k indices are then written in no particular order.
//   A row-half that collects more than KT_CAP candidates inside the bound (floods of exact ties; feature spaces so
//   collapsed that the pinned chain's own rounding exceeds the neighbour spacing) raises a flag for its 64-row tile, and
//   knn.cu's exact kernel redoes just those tiles inside the same call.
#include <cstdlib>
#include <map>
#include <mutex>

#include "common.cuh"

namespace gfs {

constexpr int KT_ROWS = 256;          // query rows per CTA
constexpr int KT_COLS = 64;           // candidates per stage
constexpr int KT_TST = 4;             // TMEM stages: 2 row tiles x 4 stages x 64 fp32 columns = 512 columns
constexpr int KT_BST = 4;             // shared-memory stages of the candidate operand
constexpr int KT_ATMEM_DEFAULT = 1;   // query operand of the filter MMAs: 0 = shared memory, 1 = tensor memory
constexpr int KT_CST = KT_BST + KT_TST;   // slots of the per-candidate constants: loaded with the operand, read after the MMA
constexpr int KT_SELW = 16;           // selection warps: (row tile, lane quarter, column half)
constexpr int KT_THREADS = 64 + 32 * KT_SELW;   // warp 0 TMA, warp 1 MMA, warps 2-17 selection
constexpr int KT_CAP = 128;           // survivor records per (row, column half)
constexpr int KT_SURV = 2 * KT_CAP;   // survivors per row the finish kernel can take (32 per round, one per lane)
constexpr int KF_RMAX = 20;           // SET finish: rows with up to k + KF_RMAX survivors are classified by their bounds
constexpr int KF_GL = 4;              // finish: row-gather load instructions in flight per group

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t kt_ord_key(float d) {
    const uint32_t u = __float_as_uint(d);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float kt_ord_val(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// ---------------------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------------------
// One CTA per 64 points (half a 128-row operand tile): 16 CTAs of 128 threads per block of 1024 points keep far more loads
// in flight than one CTA per tile did (the kernel is latency bound: it reads C x N floats per block once and writes the
// operand tiles, the point-major copy and four small per-point arrays).  The (C x 64) slab is staged in shared memory with
// coalesced loads; thread t < 64 then owns point t for the norm chains, and the operand tiles / the point-major copy are
// written with (row, 16-byte chunk) work items so that consecutive threads write consecutive bytes.
constexpr int KP_PTS = 64;        // points per CTA
constexpr int KP_LD = 65;         // padded row of the staged slab
constexpr int KP_THREADS = 128;
__global__ void __launch_bounds__(KP_THREADS)
knn_prep_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, int Npad, int Cp16, int KB, int CPT,
                uint8_t* __restrict__ ops, float* __restrict__ xp, float* __restrict__ nh, float* __restrict__ nl,
                uint32_t* __restrict__ tag, float* __restrict__ sqnorm) {
    pdl_enter();
    __shared__ float xs[64 * KP_LD];
    __shared__ float mus[64];
    const int t = threadIdx.x, b = blockIdx.y;
    const int rt = blockIdx.x >> 1, roff = (blockIdx.x & 1) * KP_PTS;      // operand tile, first row inside it
    const int n0 = blockIdx.x * KP_PTS, n = n0 + t;
    const bool valid = t < KP_PTS && n < N;
    const float* xb = x + (int64_t)b * bstride;
    // the shift: any vector is valid (distances are translation invariant), it only has to be the SAME for every point of
    // the block and close to the data.  The mean of four spread-out points costs four loads and no extra kernel.
    if (t < 64) {
        const float* p = xb + (int64_t)t * N;
        mus[t] = t < C ? 0.25f * ((__ldg(p) + __ldg(p + N / 4)) + (__ldg(p + N / 2) + __ldg(p + 3 * (N / 4)))) : 0.0f;
    }
    for (int i = t; i < Cp16 * KP_PTS; i += KP_THREADS) {
        const int c = i >> 6, r = i & 63;
        xs[c * KP_LD + r] = (n0 + r < N && c < C) ? __ldg(xb + (int64_t)c * N + n0 + r) : 0.0f;
    }
    __syncthreads();

    if (t < KP_PTS) {
    float xx = 0.0f, cc = 0.0f, aw = 0.0f;
    double xd = 0.0;
    for (int c = 0; c < C; ++c) {
        const float v = xs[c * KP_LD + t];
        const float u = v - mus[c];
        xx = fmaf(v, v, xx);                 // same chain as sqnorm_kernel
        cc = fmaf(u, u, cc);
        aw = __fmaf_ru(__fmul_ru(v, v), (float)(C - c), aw);   // A = sum_c (C - c) x_c^2, rounded up
        xd = fma((double)v, (double)v, xd);  // |x|^2 to 2^-53: what the fp32 chain above SHOULD have given
    }
    if (!valid) cc = 0.0f;                   // padding rows: all-zero operands (their v - mu would not be zero)
    // The pinned distance subtracts the COMPUTED square norm xx, not |x_j|^2: its rounding error is a known number per
    // point, not an uncertainty.  dj = (|x_j|^2 - xx_j) / 2 moves the candidate's filter constant to where the pinned
    // arithmetic will put it (the row's own term is the same for all its candidates and does not matter).
    const float dj = (float)(0.5 * (xd - (double)xx));
    // a = the point's share of the pair error bound  |v(i,j) - D(i,j)| <= a_i + a_j  (see the header):
    //   2^-15 |x~|^2  covers the bf16 split residual, the tensor core's fp32 accumulation and the fp32 addition of the constant
    //   2^-25 A + 6 2^-25 |x|^2  the rounding of the pinned fp32 dot product on the original coordinates, A = sum_c (C - c) x_c^2:
    //   channel c's product sits in C - c of the chain's partial sums, each of which is rounded once.
    // nh carries +a_j (-> an upper bound of D), nl carries -a_j (-> a lower bound), the tag 2 a_j rounded UP to bf16.
    const float a = __fmaf_ru(0x1p-15f, cc, __fmaf_ru(0x1p-25f, aw, __fmul_ru(6.0f * 0x1p-25f, xx)));
    const __nv_bfloat16 dl = __float2bfloat16_ru(2.0f * a);
    nh[(int64_t)b * Npad + n] = valid ? __fadd_ru(__fmaf_ru(-0.5f, cc, a), dj) : -INFINITY;
    nl[(int64_t)b * Npad + n] = valid ? __fadd_rd(__fmaf_rd(-0.5f, cc, -a), dj) : -INFINITY;
    tag[(int64_t)b * Npad + n] = ((uint32_t)n << 16) | (uint32_t)__bfloat16_as_ushort(dl);
    if (valid) sqnorm[(int64_t)b * N + n] = xx;
    }

    // operand tiles: columns [h: 0..Cp16) [l: Cp16..2 Cp16), 8 columns per 16-byte chunk
    uint8_t* tiles = ops + ((int64_t)b * (Npad / 128) + rt) * KB * 16384;
    const int cpr = Cp16 >> 3;               // chunks per row and per part
    for (int i = t; i < KP_PTS * cpr; i += KP_THREADS) {
        const int r = i / cpr, q = i - r * cpr;
        const bool on = n0 + r < N;
        uint32_t hp[4], lp[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = q * 8 + 2 * e;
            const float u0 = (on && c < C) ? xs[c * KP_LD + r] - mus[c] : 0.0f;
            const float u1 = (on && c + 1 < C) ? xs[(c + 1) * KP_LD + r] - mus[c + 1] : 0.0f;
            const __nv_bfloat16 h0 = __float2bfloat16_rn(u0), h1 = __float2bfloat16_rn(u1);
            __nv_bfloat162 hh;
            hh.x = h0;
            hh.y = h1;
            hp[e] = *reinterpret_cast<uint32_t*>(&hh);
            lp[e] = pack_bf16x2(u0 - __bfloat162float(h0), u1 - __bfloat162float(h1));
        }
        const int ch = q * 8, cl = Cp16 + q * 8;
        *reinterpret_cast<uint4*>(tiles + (ch >> 6) * 16384 + sw128(roff + r, (ch & 63) >> 3)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(tiles + (cl >> 6) * 16384 + sw128(roff + r, (cl & 63) >> 3)) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
    // point-major fp32 copy of the ORIGINAL coordinates, rows of CPT floats (zero padded)
    float* xrows = xp + ((int64_t)b * Npad + n0) * CPT;
    const int fpr = CPT >> 2;
    for (int i = t; i < KP_PTS * fpr; i += KP_THREADS) {
        const int r = i / fpr, q = i - r * fpr;
        float4 o;
        o.x = 4 * q < Cp16 ? xs[(4 * q) * KP_LD + r] : 0.0f;
        o.y = 4 * q + 1 < Cp16 ? xs[(4 * q + 1) * KP_LD + r] : 0.0f;
        o.z = 4 * q + 2 < Cp16 ? xs[(4 * q + 2) * KP_LD + r] : 0.0f;
        o.w = 4 * q + 3 < Cp16 ? xs[(4 * q + 3) * KP_LD + r] : 0.0f;
        *reinterpret_cast<float4*>(xrows + (int64_t)r * CPT + q * 4) = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------------------------
struct KtCtl {
    float cst[KT_CST][KT_COLS];       // the stage's per-candidate constant: nl_j in pass A, nh_j in pass B
    uint32_t tag[KT_CST][KT_COLS];    // (j << 16) | bf16(2 a_j)   (pass B only)
    uint64_t a_full, b_full[KT_BST], b_empty[KT_BST], d_full[KT_TST], d_empty[KT_TST];
    float* surv_base;
    uint32_t tmem_base;
};

// mask |= bit if v >= thr: a compare and a predicated OR
__device__ __forceinline__ void kt_mask_ge(uint32_t& mask, float v, float thr, uint32_t bit) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(mask) : "f"(v), "f"(thr), "r"(bit));
}

// bitonic sorting network on registers, descending; every index is a compile-time constant after unrolling
template <int NG>
__device__ __forceinline__ void kt_sort_desc(float (&a)[NG]) {
#pragma unroll
    for (int k = 2; k <= NG; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                const int l = i ^ j;
                if (l > i) {
                    const float hi = fmaxf(a[i], a[l]), lo = fminf(a[i], a[l]);
                    const bool desc = (i & k) == 0;
                    a[i] = desc ? hi : lo;
                    a[l] = desc ? lo : hi;
                }
            }
        }
    }
}

// ATM: the query operand is copied into tensor memory once per CTA (tcgen05.cp, 64 columns per row tile) and every MMA
// reads it from there: the shared-memory pipe then carries only the candidate operand (2 KB per MMA instead of 6 KB).  At
// C = 64 that pipe is what the kernel runs out of (ncu: tensor-core wavefronts 56 % + LSU wavefronts 29 % of its peak).
// The accumulators keep 384 of the 512 columns: three stages per row tile instead of four.
template <int NG, int KS, bool ATM>
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc_kernel(const uint8_t* __restrict__ ops, const float* __restrict__ nh, const float* __restrict__ nl,
              const uint32_t* __restrict__ tag, int* __restrict__ flags, int N, int Npad, int Cp16, int KB, int k,
              float* __restrict__ surv, int* __restrict__ surv_cnt, float* __restrict__ dbg) {
    pdl_wait();
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // [A: 2 row tiles x KB k-blocks x 16 KiB][B: KT_BST stages x KB x 8 KiB (64 candidate rows)][xch: 2 x NG/4 x 256 float4][ctl]
    unsigned char* sA = base;
    const uint32_t a_bytes = (uint32_t)KB * 16384u;   // one row tile of the query operand
    const uint32_t b_bytes = (uint32_t)KB * 8192u;    // one candidate stage
    unsigned char* sB0 = sA + 2 * a_bytes;
    float4* xch = reinterpret_cast<float4*>(sB0 + KT_BST * b_bytes);
    KtCtl& s = *reinterpret_cast<KtCtl*>(reinterpret_cast<unsigned char*>(xch) + 2 * (NG / 4) * KT_ROWS * 16);

    constexpr int TST = ATM ? 3 : KT_TST;             // accumulator stages per row tile
    constexpr uint32_t D0 = ATM ? 128u : 0u;          // first accumulator column (ATM: [A tile 0: 64][A tile 1: 64][D ...])
    constexpr uint32_t DT = TST * KT_COLS;            // accumulator columns per row tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * KT_ROWS;
    const int nst = Npad / KT_COLS;
    const uint8_t* blk = ops + (int64_t)b * (Npad / 128) * a_bytes;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&s.a_full, 1);
            for (int i = 0; i < KT_BST; ++i) {
                mbar_init(&s.b_full[i], 1);
                mbar_init(&s.b_empty[i], 1);
            }
            for (int i = 0; i < KT_TST; ++i) {
                mbar_init(&s.d_full[i], 1);
                mbar_init(&s.d_empty[i], KT_SELW);
            }
            mbar_fence_init();
            s.surv_base = surv;
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
            mbar_arrive_expect_tx(&s.a_full, 2u * a_bytes);
            tma_load_1d(sA, blk + (int64_t)(q0 / 128) * a_bytes, 2u * a_bytes, &s.a_full);
            for (int st = 0; st < 2 * nst; ++st) {
                const bool passB = st >= nst;
                const int cs = passB ? st - nst : st;                 // candidate stage
                const int buf = st % KT_BST, slot = st % KT_CST;
                // The operand buffer is free once the MMAs of stage st - KT_BST have read it.  Those MMAs had to wait for the
                // selection warps to hand back the accumulator of stage st - KT_BST - KT_TST, so the constants' slot
                // (st mod KT_CST, last used by that stage) is free as well: the loads never wait for the selection warps
                // themselves, only the MMAs do, and the operand is in shared memory by the time an accumulator frees up.
                mbar_wait(&s.b_empty[buf], ((st / KT_BST) & 1) ^ 1);
                mbar_arrive_expect_tx(&s.b_full[buf], b_bytes + KT_COLS * (passB ? 8u : 4u));
                // candidates [cs*64, cs*64+64): half of row tile cs/2, every k-block
                const uint8_t* src = blk + (int64_t)(cs >> 1) * a_bytes + (cs & 1) * 8192;
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_1d(sB0 + buf * b_bytes + kb * 8192, src + (int64_t)kb * 16384, 8192u, &s.b_full[buf]);
                tma_load_1d(s.cst[slot], (passB ? nh : nl) + (int64_t)b * Npad + cs * KT_COLS, KT_COLS * 4u, &s.b_full[buf]);
                if (passB) tma_load_1d(s.tag[slot], tag + (int64_t)b * Npad + cs * KT_COLS, KT_COLS * 4u, &s.b_full[buf]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        // The whole warp runs the loop (uniform control flow, uniform values: the descriptors live in uniform registers) and
        // one elected lane issues; KS (16-column k-steps of one operand part) is a template parameter so that the 6 KS
        // tcgen05.mma of a stage are straight-line code with immediate descriptor offsets.
        const uint32_t idesc = umma_idesc_bf16(128, KT_COLS);
        constexpr int CP16 = KS * 16;
        mbar_wait(&s.a_full, 0);
        const uint64_t a0 = umma_desc_sw128(smem_u32(sA));
        const uint64_t b0 = umma_desc_sw128(smem_u32(sB0));
        const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4;     // descriptor address units (16 bytes)
        if (ATM) {
            // the query operand [h | l] of both row tiles -> tensor memory, 16 columns (8 TMEM columns) per copy
            if (elect_one()) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int q = 0; q < 2 * KS; ++q) {
                        const int c16 = q * 16;
                        const uint32_t oa = (uint32_t)((c16 >> 6) * 16384 + (c16 & 63) * 2) >> 4;
                        tmem_cp_128x256b(tmem + mt * 64 + q * 8, a0 + (uint64_t)(mt * a_step) + oa);
                    }
            }
            __syncwarp();
        }
        for (int st = 0; st < 2 * nst; ++st) {
            const int buf = st % KT_BST, ts = st % TST;
            mbar_wait(&s.d_empty[ts], ((st / TST) & 1) ^ 1);
            mbar_wait(&s.b_full[buf], (st / KT_BST) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t bB = b0 + (uint64_t)(buf * b_step);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const uint64_t aB = a0 + (uint64_t)(mt * a_step);
                    const uint32_t d = tmem + D0 + mt * DT + ts * KT_COLS;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const int ch = ks * 16, cl = CP16 + ks * 16;
                        // offset of a 16-column slice inside the operand: k-block (16 KiB / 8 KiB apart) + 32 bytes per k-step
                        const uint32_t oah = (uint32_t)((ch >> 6) * 16384 + (ch & 63) * 2) >> 4, oal = (uint32_t)((cl >> 6) * 16384 + (cl & 63) * 2) >> 4;
                        const uint32_t obh = (uint32_t)((ch >> 6) * 8192 + (ch & 63) * 2) >> 4, obl = (uint32_t)((cl >> 6) * 8192 + (cl & 63) * 2) >> 4;
                        if (ATM) {
                            const uint32_t ah = tmem + mt * 64 + ks * 8, al = tmem + mt * 64 + (KS + ks) * 8;
                            umma_bf16_ts(d, ah, bB + obh, idesc, ks ? 1u : 0u);
                            umma_bf16_ts(d, ah, bB + obl, idesc, 1u);
                            umma_bf16_ts(d, al, bB + obh, idesc, 1u);
                        } else {
                            umma_bf16(d, aB + oah, bB + obh, idesc, ks ? 1u : 0u);
                            umma_bf16(d, aB + oah, bB + obl, idesc, 1u);
                            umma_bf16(d, aB + oal, bB + obh, idesc, 1u);
                        }
                    }
                }
                umma_commit(&s.b_empty[buf]);
                umma_commit(&s.d_full[ts]);
            }
            __syncwarp();
        }
    } else {
        // =============================== selection: two threads per query row ===============================
        // A warp may only read the TMEM lanes 32 (warp % 4) .. +31.  Among the 16 warps every residue occurs four times:
        // they take (row tile, column half) = (0,0) (1,0) (0,1) (1,1) in warp order.
        const int quarter = warp & 3, slot = (warp - 2) >> 2;
        const int mt = slot & 1, half = slot >> 1;
        const int row = mt * 128 + quarter * 32 + lane;
        const int n = q0 + row;
        const bool valid = n < N;
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + D0 + mt * DT + half * 32;
        uint32_t r[32];

        // ---------------- pass A: running maxima of the lower bounds, one group per column position ----------------
        float gm[NG];
#pragma unroll
        for (int i = 0; i < NG; ++i) gm[i] = -INFINITY;
        for (int st = 0; st < nst; ++st) {
            const int ts = st % TST;
            mbar_wait(&s.d_full[ts], (st / TST) & 1);
            tc_fence_after();
            tmem_ld32(tbase + ts * KT_COLS, r);
            tmem_ld_wait32(r);
            const float* cst = s.cst[st % KT_CST] + half * 32;
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(cst + c);
                const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), make_float2(l4.x, l4.y));
                const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])), make_float2(l4.z, l4.w));
                gm[c % NG] = fmaxf(gm[c % NG], s0.x);
                gm[(c + 1) % NG] = fmaxf(gm[(c + 1) % NG], s0.y);
                gm[(c + 2) % NG] = fmaxf(gm[(c + 2) % NG], s1.x);
                gm[(c + 3) % NG] = fmaxf(gm[(c + 3) % NG], s1.y);
            }
            // hand the TMEM stage (and its constants' slot) back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.d_empty[ts]);
        }

        // ---------------- the row's threshold: k-th largest of the 2 NG group maxima ----------------
        kt_sort_desc<NG>(gm);
#pragma unroll
        for (int q = 0; q < NG / 4; ++q)
            xch[(half * (NG / 4) + q) * KT_ROWS + row] = make_float4(gm[4 * q], gm[4 * q + 1], gm[4 * q + 2], gm[4 * q + 3]);
        named_bar_sync(1, 32 * KT_SELW);
        float thr;
        {
            const float* xf = reinterpret_cast<const float*>(xch);
            // g-th largest (0-based) of this half (mine) / of the other half
            auto mine = [&](int g) { return xf[((half * (NG / 4) + (g >> 2)) * KT_ROWS + row) * 4 + (g & 3)]; };
            auto other = [&](int g) { return xf[(((half ^ 1) * (NG / 4) + (g >> 2)) * KT_ROWS + row) * 4 + (g & 3)]; };
            // take i from mine and k - i from other: the smallest i with other[k-i-1] >= mine[i]
            int lo = k > NG ? k - NG : 0, hi = k < NG ? k : NG;
            while (lo < hi) {
                const int i = (lo + hi) >> 1;
                if (other(k - i - 1) < mine(i)) lo = i + 1;
                else hi = i;
            }
            const float ta = lo > 0 ? mine(lo - 1) : INFINITY;
            const float tb = k - lo > 0 ? other(k - lo - 1) : INFINITY;
            const float T = fminf(ta, tb);
            // margin = 2 a_i (the row's own share of the pair error bound, taken from its tag: rounded up).
            // -FLT_MAX, not -inf: padded candidates carry -inf and must never pass, not even while the bound is unknown
            const float margin = valid ? __uint_as_float(tag[(int64_t)b * Npad + n] << 16) : 0.0f;
            thr = valid ? fmaxf(__fsub_rd(T, margin), -3.402823466e38f) : INFINITY;
        }

        // ---------------- pass B: append every candidate whose upper bound reaches the threshold ----------------
        // The scan is two instructions per candidate (compare, predicated OR into a bit mask) next to the packed add; the few
        // hits (~1 % of the candidates) are then peeled off the mask and written as {u, tag} records.  The filter values of
        // the chunk are staged in the thread's own shared-memory slot (the exchange area of the threshold step, now free)
        // because a hit's value has to be fetched by a run-time index.
        // 32-bit record offsets from the (uniform) base pointer: one IMAD.WIDE per store instead of 64-bit pointer chains
        // (records are {u, tag} pairs, the two half lists of a row lie side by side: [half 0: KT_CAP x 8 B][half 1: KT_CAP x 8 B])
        const int64_t g = (int64_t)b * N + (valid ? n : 0);
        const uint32_t off0 = (uint32_t)(g * 2 + half) * (uint32_t)KT_CAP;           // KT_CAP records {u, tag} of 8 bytes
        uint32_t off = off0;
        const uint32_t olim = off0 + (KT_CAP - 32);
        // the base comes from shared memory, not from the parameter bank: it then lives in a register pair instead of being
        // re-loaded from the constant bank for every record
        uint2* const survw = reinterpret_cast<uint2*>(s.surv_base);
        bool ovf = false;
        named_bar_sync(1, 32 * KT_SELW);                          // every thread has read the other half's maxima: xch is free
        float4* const stg = xch + (tid - 64);                     // float4 q of this thread's chunk lives at stg[q * 512]
        const float* const stgf = reinterpret_cast<const float*>(stg);
        for (int st = nst; st < 2 * nst; ++st) {
            const int ts = st % TST;
            mbar_wait(&s.d_full[ts], (st / TST) & 1);
            tc_fence_after();
            tmem_ld32(tbase + ts * KT_COLS, r);
            tmem_ld_wait32(r);
            if (off > olim) {         // fewer than 32 free records: this row-half is left to the exact repair pass
                ovf = true;
                thr = INFINITY;
            }
            const float* cst = s.cst[st % KT_CST] + half * 32;
            const uint32_t* tgs = s.tag[st % KT_CST] + half * 32;
            uint32_t m0 = 0u, m1 = 0u;                            // two chains: even / odd pairs
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
                const float4 h4 = *reinterpret_cast<const float4*>(cst + c);
                const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), make_float2(h4.x, h4.y));
                const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])), make_float2(h4.z, h4.w));
                stg[(c >> 2) * (32 * KT_SELW)] = make_float4(s0.x, s0.y, s1.x, s1.y);
                kt_mask_ge(m0, s0.x, thr, 1u << c);
                kt_mask_ge(m0, s0.y, thr, 1u << (c + 1));
                kt_mask_ge(m1, s1.x, thr, 1u << (c + 2));
                kt_mask_ge(m1, s1.y, thr, 1u << (c + 3));
            }
            if (dbg && valid) {   // diagnostic entry only: dump the filter value u = D' - |x~_j|^2/2 + a_j
                float* o = dbg + ((int64_t)b * N + n) * Npad + (st - nst) * KT_COLS + half * 32;
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] = stgf[(c >> 2) * (128 * KT_SELW) + (c & 3)];
            }
            uint32_t mask = m0 | m1;
            while (mask) {                                        // ascending column order
                const int c = __ffs(mask) - 1;
                mask &= mask - 1;
                survw[off] = make_uint2(__float_as_uint(stgf[(c >> 2) * (128 * KT_SELW) + (c & 3)]), tgs[c]);   // one 8-byte store
                ++off;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.d_empty[ts]);
        }
        if (valid) {
            surv_cnt[g * 2 + half] = ovf ? -1 : (int)(off - off0);
            if (ovf) flags[(int64_t)b * ((N + 63) / 64) + n / 64] = 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// finish: exact pinned distances of the survivors.  One warp per query row, one lane per survivor.
// The survivors' fp32 rows are gathered with coalesced loads (2 or 8 rows per instruction) into a padded staging tile,
// each lane then runs the pinned fma chain over its own row; ranking by counting over 64-bit keys
// (orderable d << 32 | ~j) puts the k best first: nearest first, ties -> ascending index.
//
// SET mode: the caller only needs the neighbour SET (the fused EdgeConv takes a max over it).  With
//   D_j in [w_j - a_i, u_j + a_i],  w_j = u_j - 2 a_j,  m = 2 a_i
// a survivor is certainly among the k best if fewer than k candidates can possibly reach it (w_j > U + m, U = the (k+1)-th
// largest u of the row), certainly not if k candidates certainly beat it (u_j + m < W, W = the k-th largest w); candidates
// the filter dropped have u < T - m <= W - m and change neither count.  Only the band in between gets the pinned
// arithmetic, and the best (k - #certain) of the band complete the set.  The k indices are written in no particular order.
// ---------------------------------------------------------------------------------------------------------------
template <int CPT, int KF_WARPS, bool SET>
__global__ void __launch_bounds__(KF_WARPS * 32, CPT == 16 ? 5 : (SET ? 7 : 4))
knn_finish_kernel(const float* __restrict__ xp, const float* __restrict__ sqnorm, const float* __restrict__ surv,
                  const int* __restrict__ surv_cnt, const uint32_t* __restrict__ tag, int N, int Npad, int k, int64_t rows,
                  int32_t* __restrict__ idx_out, float* __restrict__ dist_out) {
    pdl_wait();
    constexpr int RS = CPT + 4;            // padded staging row: conflict-free LDS.128 across lanes
    constexpr int LPR = CPT / 4;           // lanes that fetch one row
    constexpr int RPI = 32 / LPR;          // rows per load instruction
    // candidates per round of exact arithmetic (one lane each).  SET mode leaves a handful per row (and none at all for most
    // rows), so its staging tile is half as high: the kernel waits for loads and lives on the number of resident warps.
    constexpr int RND = (SET && CPT == 64) ? 16 : 32;
    __shared__ __align__(16) float stg_all[KF_WARPS][RND * RS];
    __shared__ __align__(16) unsigned long long kbuf[KF_WARPS][KT_SURV];
    __shared__ __align__(16) float xi_all[KF_WARPS][CPT];
    __shared__ uint16_t jbuf[KF_WARPS][KT_SURV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stg = stg_all[warp];
    const int sub = lane % LPR, grp = lane / LPR;
    const int64_t gstep = (int64_t)gridDim.x * KF_WARPS;
    const uint2* const recs = reinterpret_cast<const uint2*>(surv);      // [row][half][KT_CAP] records {u, tag}
    int64_t g = (int64_t)blockIdx.x * KF_WARPS + warp;
    // SET mode keeps the NEXT row's counters and the first 32 records of either half list in registers: one round of loads
    // issued a whole row ahead, instead of two dependent rounds (counters, then records) at the head of every row -- the
    // kernel spends its time waiting for exactly these loads.  Slots beyond a list's count hold whatever the workspace
    // held; they are masked by the counters.
    int2 pc = make_int2(0, 0);
    uint2 pr0 = make_uint2(0u, 0u), pr1 = make_uint2(0u, 0u);
    uint32_t pm = 0u;
    if (SET && g < rows) {
        pc = __ldg(reinterpret_cast<const int2*>(surv_cnt) + g);
        pr0 = __ldg(recs + g * (2 * KT_CAP) + lane);
        pr1 = __ldg(recs + g * (2 * KT_CAP) + KT_CAP + lane);
        pm = __ldg(tag + (g / N) * Npad + g % N);
    }
    for (; g < rows; g += gstep) {
        int c0, c1;
        uint2 r0 = pr0, r1 = pr1;
        uint32_t mtag = pm;
        if (SET) {
            c0 = pc.x;
            c1 = pc.y;
            const int64_t gn = g + gstep;
            if (gn < rows) {
                pc = __ldg(reinterpret_cast<const int2*>(surv_cnt) + gn);
                pr0 = __ldg(recs + gn * (2 * KT_CAP) + lane);
                pr1 = __ldg(recs + gn * (2 * KT_CAP) + KT_CAP + lane);
                pm = __ldg(tag + (gn / N) * Npad + gn % N);
            }
        } else {
            c0 = __ldg(surv_cnt + 2 * g);
            c1 = __ldg(surv_cnt + 2 * g + 1);
        }
        if (c0 < 0 || c1 < 0) continue;              // flagged for the exact repair pass
        const int ns = c0 + c1;
        const uint2* su = recs + g * (2 * KT_CAP);   // [half 0][half 1]
        const int64_t b = g / N;
        const float* xpb = xp + b * Npad * CPT;
        u64* keys = reinterpret_cast<u64*>(kbuf[warp]);
        uint16_t* js = jbuf[warp];
        int nb = ns;                                 // candidates that need the exact arithmetic
        int kneed = k;                               // ... of which this many complete the set
        int nout = 0;                                // (SET) indices already written
        if (SET) {
            // ---- classify by the bounds.  Slots: lane of half list 0, lane of half list 1 ----
            const int r = ns - k;                    // survivors that have to go
            if (c0 <= 32 && c1 <= 32 && r <= KF_RMAX) {
                float u[2], w[2];
                uint32_t tg[2], ku[2];
                float amax = 0.0f;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int e = h * 32 + lane;
                    const bool live = lane < (h ? c1 : c0);
                    const uint2 rec = h ? r1 : r0;
                    u[h] = w[h] = INFINITY;
                    tg[h] = 0u;
                    if (live) {
                        u[h] = __uint_as_float(rec.x);
                        tg[h] = rec.y;
                        const float a2 = __uint_as_float(tg[h] << 16);
                        w[h] = __fsub_rd(u[h], a2);
                        amax = fmaxf(amax, a2);
                    }
                    // unique keys: the low 6 bits carry the slot, so every REDUX below removes exactly one survivor.  Truncation
                    // is monotone: the i-th smallest truncated key is the truncation of the i-th smallest u (empty slots: +inf)
                    ku[h] = (kt_ord_key(u[h]) & ~63u) | (uint32_t)e;
                }
                // U1 = (k+1)-th largest u = r-th smallest, U0 = k-th largest u = (r+1)-th smallest: r is small (the filter keeps
                // k plus a handful), so the smallest keys are peeled off one warp-wide integer min (REDUX) at a time.
                uint32_t k1 = 0u, k0 = 0u;           // key 0 sorts below every float: "no such element" = -inf
                for (int it = 0; it <= r; ++it) {
                    const uint32_t mn = __reduce_min_sync(0xffffffffu, min(ku[0], ku[1]));
                    k1 = k0;
                    k0 = mn;
                    ku[0] = ku[0] == mn ? 0xffffffffu : ku[0];
                    ku[1] = ku[1] == mn ? 0xffffffffu : ku[1];
                }
                // U: an upper bound of the (k+1)-th largest u (low bits filled); W: a lower bound of the k-th largest lower
                // bound w (w_j >= u_j - max_j 2 a_j, low bits cleared)
                const float U = r > 0 ? kt_ord_val(k1 | 63u) : -INFINITY;
                amax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(amax)));     // non-negative floats order as integers
                const float W = __fsub_rd(kt_ord_val(k0 & ~63u), amax);
                const float m = __uint_as_float(mtag << 16);
                const float Uin = __fadd_ru(U, m), Wout = __fsub_rd(W, m);
                int nin = 0;
                nb = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool live = lane < (h ? c1 : c0);
                    const bool in = live && w[h] > Uin;                       // at most k - 1 others can reach it
                    const bool out = live && !in && u[h] < Wout;              // k others certainly beat it
                    const bool band = live && !in && !out;
                    const unsigned bi = __ballot_sync(0xffffffffu, in), bb = __ballot_sync(0xffffffffu, band);
                    const unsigned lt = (1u << lane) - 1u;
                    if (in) idx_out[g * k + nin + __popc(bi & lt)] = (int32_t)(tg[h] >> 16);
                    if (band) js[nb + __popc(bb & lt)] = (uint16_t)(tg[h] >> 16);
                    nin += __popc(bi);
                    nb += __popc(bb);
                }
                nout = nin;
                kneed = k - nin;
                __syncwarp();
            } else {
                for (int e = lane; e < ns; e += 32) js[e] = (uint16_t)(__ldg(&(e < c0 ? su + e : su + KT_CAP + (e - c0))->y) >> 16);
                __syncwarp();
            }
            if (kneed == 0) continue;
        } else {
            for (int e = lane; e < ns; e += 32) js[e] = (uint16_t)(__ldg(&(e < c0 ? su + e : su + KT_CAP + (e - c0))->y) >> 16);
            __syncwarp();
        }
        // the query row itself goes to shared memory (broadcast reads in the chain): holding it in 64 registers per lane
        // would halve the number of resident warps, and this kernel lives on latency hiding
        float* xis = xi_all[warp];
        if (lane < LPR) *reinterpret_cast<float4*>(xis + lane * 4) = __ldg(reinterpret_cast<const float4*>(xpb + (g - b * N) * CPT + lane * 4));
        const float xxi = __ldg(sqnorm + g);
        for (int base = 0; base < nb; base += RND) {   // further rounds only for rows with more than RND candidates
            const int e = base + lane;
            const int nbr = nb - base;                 // candidates of this round: the first min(nbr, RND) lanes
            const bool act = lane < RND && e < nb;
            const uint32_t j = act ? (uint32_t)js[e] : 0u;
            const float xxj = __ldg(sqnorm + b * N + j);
            __syncwarp();                            // the previous chains are done with the staging tile
            // gather in groups of KF_GL load instructions (uniform test: groups beyond the round's candidates are skipped;
            // the band is a handful of candidates for most rows)
#pragma unroll
            for (int g0 = 0; g0 < RND / RPI; g0 += KF_GL) {
                if (g0 * RPI < nbr) {
                    float4 t[KF_GL];
#pragma unroll
                    for (int it = 0; it < KF_GL; ++it) {
                        const uint32_t jj = __shfl_sync(0xffffffffu, j, (g0 + it) * RPI + grp);
                        t[it] = __ldg(reinterpret_cast<const float4*>(xpb + (int64_t)jj * CPT + sub * 4));
                    }
#pragma unroll
                    for (int it = 0; it < KF_GL; ++it) *reinterpret_cast<float4*>(stg + ((g0 + it) * RPI + grp) * RS + sub * 4) = t[it];
                }
            }
            __syncwarp();
            float dot = 0.0f;
#pragma unroll
            for (int c = 0; c < CPT; c += 4) {
                const float4 q = *reinterpret_cast<const float4*>(stg + (lane & (RND - 1)) * RS + c);
                const float4 a = *reinterpret_cast<const float4*>(xis + c);
                dot = fmaf(a.x, q.x, dot);
                dot = fmaf(a.y, q.y, dot);
                dot = fmaf(a.z, q.z, dot);
                dot = fmaf(a.w, q.w, dot);
            }
            const float d = fmaf(2.0f, dot, -xxi) - xxj;
            // 64-bit key: larger = nearer, equal distances -> smaller index first; 0 = empty slot (below every real key)
            if (lane < RND) keys[e] = act ? (((u64)kt_ord_key(d) << 32) | (u64)(~j)) : 0ull;
        }
        __syncwarp();
        // rank by counting (independent broadcast reads: no dependent shuffle network), keys are all distinct
        const int nk = (nb + 3) & ~3;                  // slots up to the next multiple of RND hold 0 = never greater
        for (int half = 0; half * 32 < nb; ++half) {
            const u64 mine = half * 32 + lane < nb ? keys[half * 32 + lane] : 0ull;
            int rank = 0;
#pragma unroll 2
            for (int f = 0; f < nk; f += 4) {            // two 16-byte broadcast loads = four keys
                const ulonglong2 q0 = *reinterpret_cast<const ulonglong2*>(keys + f);
                const ulonglong2 q1 = *reinterpret_cast<const ulonglong2*>(keys + f + 2);
                if (q0.x > mine) ++rank;
                if (q0.y > mine) ++rank;
                if (q1.x > mine) ++rank;
                if (q1.y > mine) ++rank;
            }
            if (mine != 0ull && rank < kneed) {
                idx_out[g * k + nout + rank] = (int32_t)(~(uint32_t)mine);
                if (!SET && dist_out) dist_out[g * k + rank] = kt_ord_val((uint32_t)(mine >> 32));
            }
        }
        __syncwarp();                                // keys / js are reused by the next row
    }
}

struct KtPlan {
    int Npad, Cp16, KB, CPT;
    size_t off_xp, off_nh, off_nl, off_tag, off_surv, off_cnt, off_flags, total, zero_bytes;
};
static KtPlan kt_plan(int B, int C, int N) {
    KtPlan p;
    p.Npad = (N + KT_ROWS - 1) / KT_ROWS * KT_ROWS;
    p.Cp16 = (C + 15) / 16 * 16;
    p.CPT = C <= 16 ? 16 : 64;
    p.KB = (2 * p.Cp16 + 63) / 64;
    size_t o = (size_t)B * (p.Npad / 128) * p.KB * 16384;
    p.off_xp = o;
    o += (size_t)B * p.Npad * p.CPT * 4;
    p.off_nh = o;
    o += (size_t)B * p.Npad * 4;
    p.off_nl = o;
    o += (size_t)B * p.Npad * 4;
    p.off_tag = o;
    o += (size_t)B * p.Npad * 4;
    p.off_surv = o;
    o += (size_t)B * N * 4 * KT_CAP * 4;
    p.off_cnt = o;
    o += (size_t)B * N * 2 * 4;
    o = (o + 15) / 16 * 16;
    p.off_flags = o;
    o += (size_t)B * ((N + 63) / 64) * 4;
    p.zero_bytes = o - p.off_flags;
    p.total = (o + 255) / 256 * 256;
    return p;
}

// knn.cu: the exact kernel restricted to the flagged 64-row tiles
int knn_exact_flagged(const float* x, int64_t x_bstride, int B, int C, int N, int k, const float* sqnorm, const int* flags,
                      int32_t* idx_out, float* dist_out, cudaStream_t st);

static int g_kt_atmem = -1;            // -1: read GFS3D_KNN_ATMEM on first use; 0 / 1: query operand in shared / tensor memory
static bool kt_atmem() {
    if (g_kt_atmem < 0) {
        const char* e = getenv("GFS3D_KNN_ATMEM");
        g_kt_atmem = e ? (atoi(e) != 0) : KT_ATMEM_DEFAULT;
    }
    return g_kt_atmem != 0;
}

template <int NG, int KS, bool ATM>
static int kt_launch_filter_ks(const KtPlan& p, uint8_t* ws, int B, int N, int k, float* dbg, cudaStream_t st) {
    const size_t smem = (size_t)2 * p.KB * 16384 + (size_t)KT_BST * p.KB * 8192 + (size_t)2 * (NG / 4) * KT_ROWS * 16 + sizeof(KtCtl) + 1024;
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(knn_tc_kernel<NG, KS, ATM>), smem));
    launch_pdl<2>(knn_tc_kernel<NG, KS, ATM>, dim3(p.Npad / KT_ROWS, B), dim3(KT_THREADS), smem, st,
        ws, reinterpret_cast<const float*>(ws + p.off_nh), reinterpret_cast<const float*>(ws + p.off_nl),
        reinterpret_cast<const uint32_t*>(ws + p.off_tag), reinterpret_cast<int*>(ws + p.off_flags), N, p.Npad, p.Cp16, p.KB, k,
        reinterpret_cast<float*>(ws + p.off_surv), reinterpret_cast<int*>(ws + p.off_cnt), dbg);
    GFS_LAUNCH_OK("knn_tc_kernel");
    return GFS_OK;
}
template <int NG>
static int kt_launch_filter(const KtPlan& p, uint8_t* ws, int B, int N, int k, float* dbg, cudaStream_t st) {
    if (kt_atmem()) {
        switch (p.Cp16 >> 4) {
            case 1: return kt_launch_filter_ks<NG, 1, true>(p, ws, B, N, k, dbg, st);
            case 2: return kt_launch_filter_ks<NG, 2, true>(p, ws, B, N, k, dbg, st);
            case 3: return kt_launch_filter_ks<NG, 3, true>(p, ws, B, N, k, dbg, st);
            default: return kt_launch_filter_ks<NG, 4, true>(p, ws, B, N, k, dbg, st);
        }
    }
    switch (p.Cp16 >> 4) {
        case 1: return kt_launch_filter_ks<NG, 1, false>(p, ws, B, N, k, dbg, st);
        case 2: return kt_launch_filter_ks<NG, 2, false>(p, ws, B, N, k, dbg, st);
        case 3: return kt_launch_filter_ks<NG, 3, false>(p, ws, B, N, k, dbg, st);
        default: return kt_launch_filter_ks<NG, 4, false>(p, ws, B, N, k, dbg, st);
    }
}

template <bool SET>
static int kt_launch_finish(const KtPlan& p, uint8_t* ws, const float* sqnorm, int B, int N, int k, int32_t* idx_out, float* dist_out,
                            cudaStream_t st) {
    const int64_t rows = (int64_t)B * N;
    const int sms = sm_count();
    GFS_REQUIRE(sms > 0, GFS_ERR_CUDA, "gfs_knn_tc_f32: cannot query the device");
    const float* xp = reinterpret_cast<const float*>(ws + p.off_xp);
    const float* surv = reinterpret_cast<const float*>(ws + p.off_surv);
    const int* cnt = reinterpret_cast<const int*>(ws + p.off_cnt);
    const uint32_t* tag = reinterpret_cast<const uint32_t*>(ws + p.off_tag);
    if (p.CPT == 16) {
        const int64_t want = (rows + 7) / 8;
        const int grid = (int)(want < (int64_t)sms * 5 ? want : (int64_t)sms * 5);        // the resident CTAs, each strides over the rows
        launch_pdl<2>(knn_finish_kernel<16, 8, SET>, grid, dim3(256), 0, st,
        xp, sqnorm, surv, cnt, tag, N, p.Npad, k, rows, idx_out, dist_out);
    } else {
        const int64_t want = (rows + 3) / 4;
        const int per_sm = SET ? 7 : 4;
        const int grid = (int)(want < (int64_t)sms * per_sm ? want : (int64_t)sms * per_sm);
        launch_pdl<2>(knn_finish_kernel<64, 4, SET>, grid, dim3(128), 0, st,
        xp, sqnorm, surv, cnt, tag, N, p.Npad, k, rows, idx_out, dist_out);
    }
    GFS_LAUNCH_OK("knn_finish_kernel");
    return GFS_OK;
}

}  // namespace gfs

namespace gfs {

// One chain of the four launches over `B` consecutive blocks, scratch at `ws` (kt_plan(B, C, N) bytes).
static int kt_chain(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, uint8_t* ws, int32_t* idx_out,
                    float* dist_out, float* dbg, bool set_only, cudaStream_t st) {
    const KtPlan p = kt_plan(B, C, N);
    GFS_CUDA_OK(cudaMemsetAsync(ws + p.off_flags, 0, p.zero_bytes, st));
    launch_pdl<2>(knn_prep_kernel, dim3(p.Npad / KP_PTS, B), dim3(KP_THREADS), 0, st,
        x, x_bstride, C, N, p.Npad, p.Cp16, p.KB, p.CPT, ws,
                                                           reinterpret_cast<float*>(ws + p.off_xp),
                                                           reinterpret_cast<float*>(ws + p.off_nh),
                                                           reinterpret_cast<float*>(ws + p.off_nl),
                                                           reinterpret_cast<uint32_t*>(ws + p.off_tag), sqnorm);
    GFS_LAUNCH_OK("knn_prep_kernel");
    int rc = kt_launch_filter<32>(p, ws, B, N, k, dbg, st);
    if (rc != GFS_OK) return rc;
    rc = set_only ? kt_launch_finish<true>(p, ws, sqnorm, B, N, k, idx_out, nullptr, st)
                  : kt_launch_finish<false>(p, ws, sqnorm, B, N, k, idx_out, dist_out, st);
    if (rc != GFS_OK) return rc;
    return knn_exact_flagged(x, x_bstride, B, C, N, k, sqnorm, reinterpret_cast<const int*>(ws + p.off_flags), idx_out, dist_out, st);
}

// Two chains side by side.  The filter kernel holds one CTA per SM (all 512 TMEM columns, ~200 KB of shared memory), so a
// call whose CTA count is not a multiple of the SM count ends in a partly empty wave (32 blocks x 2048 points: 256 CTAs on
// 148 SMs, 40 SMs idle for the second half of the kernel).  Splitting the blocks into two independent chains on two
// streams lets the first chain's finish kernel (and the second chain's preparation) run on the SMs the second chain's
// filter leaves idle.  The side stream is forked from and joined back into the caller's stream with events, so the call
// keeps its stream semantics and can be captured into a CUDA graph (the graph gets two parallel branches).
struct KtSide {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static std::mutex g_kt_mu;
static std::map<int, KtSide> g_kt_side;
static int g_kt_split = -2;            // -2: read GFS3D_KNN_SPLIT on first use; -1: automatic; 0: never; n > 0: blocks in the first chain

// blocks that go to the first chain (0 = one chain)
static int kt_first_chain(int B, int N) {
    if (g_kt_split == -2) {
        const char* e = getenv("GFS3D_KNN_SPLIT");
        g_kt_split = e ? atoi(e) : -1;
        if (g_kt_split < -1) g_kt_split = -1;
    }
    if (B < 2 || g_kt_split == 0) return 0;
    if (g_kt_split > 0) return g_kt_split < B ? g_kt_split : B - 1;
    const int sms = sm_count();
    const int per_block = (N + KT_ROWS - 1) / KT_ROWS;
    const int ctas = B * per_block;
    if (sms <= 0 || ctas <= sms || ctas % sms == 0) return 0;      // one wave, or no tail to fill
    return B / 2;
}

static int kt_side_for_device(KtSide*& out) {
    int dev = 0;
    GFS_CUDA_OK(cudaGetDevice(&dev));
    KtSide& s = g_kt_side[dev];
    if (!s.stream) {
        GFS_CUDA_OK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        GFS_CUDA_OK(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
        GFS_CUDA_OK(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
    }
    out = &s;
    return GFS_OK;
}

}  // namespace gfs

extern "C" int64_t gfs_knn_tc_workspace_bytes(int B, int C, int N) {
    if (B <= 0 || C <= 0 || N <= 0 || C > 64) return 0;
    return (int64_t)gfs::kt_plan(B, C, N).total + 512;     // + the rounding slack of two chains' separate plans
}

extern "C" int gfs_knn_tc_chains(int B, int C, int N) {
    if (B <= 0 || C <= 0 || N <= 0) return 0;
    std::lock_guard<std::mutex> lk(gfs::g_kt_mu);
    return gfs::kt_first_chain(B, N) > 0 ? 2 : 1;
}

extern "C" int gfs_knn_tc_set_split(int blocks_in_first_chain) {
    std::lock_guard<std::mutex> lk(gfs::g_kt_mu);
    gfs::g_kt_split = blocks_in_first_chain < -1 ? -1 : blocks_in_first_chain;
    return GFS_OK;
}

static int kt_run(const char* who, const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, void* workspace,
                  int64_t workspace_bytes, int32_t* idx_out, float* dist_out, float* dbg, bool set_only, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x && sqnorm && idx_out && workspace, GFS_ERR_BAD_ARG, "%s: null pointer", who);
    GFS_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, GFS_ERR_BAD_ARG, "%s: non-positive size (B=%d C=%d N=%d k=%d)", who, B, C, N, k);
    GFS_REQUIRE(k <= N, GFS_ERR_BAD_ARG, "%s: k=%d exceeds N=%d", who, k, N);
    GFS_REQUIRE(k <= 40, GFS_ERR_UNSUPPORTED, "%s: k=%d > 40 is not built (use gfs_knn_f32)", who, k);
    GFS_REQUIRE(C <= 64, GFS_ERR_UNSUPPORTED, "%s: C=%d > 64 is not built", who, C);
    GFS_REQUIRE(N <= 65535, GFS_ERR_UNSUPPORTED, "%s: N=%d > 65535 (16-bit candidate slots)", who, N);
    GFS_REQUIRE((int64_t)B * N * 4 * KT_CAP < (int64_t)1 << 32, GFS_ERR_UNSUPPORTED, "%s: B*N=%lld rows exceed the 32-bit survivor offsets", who,
                (long long)B * N);
    GFS_REQUIRE(N % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, GFS_ERR_UNSUPPORTED,
                "%s: needs N %% 4 == 0 and 16-byte aligned rows (N=%d)", who, N);
    const int64_t need = gfs_knn_tc_workspace_bytes(B, C, N);
    GFS_REQUIRE(workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, GFS_ERR_BAD_ARG,
                "%s: workspace of %lld bytes (256-byte aligned) needed, got %lld", who, (long long)need, (long long)workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    std::lock_guard<std::mutex> lk(g_kt_mu);           // the side stream's events are shared by the callers of one device
    const int b1 = dbg ? 0 : kt_first_chain(B, N);     // the diagnostic entry reads one plan's flags: one chain
    if (b1 == 0) return kt_chain(x, x_bstride, B, C, N, k, sqnorm, ws, idx_out, dist_out, dbg, set_only, st);

    KtSide* side = nullptr;
    int rc = kt_side_for_device(side);
    if (rc != GFS_OK) return rc;
    const int b2 = B - b1;
    GFS_CUDA_OK(cudaEventRecord(side->fork, st));
    GFS_CUDA_OK(cudaStreamWaitEvent(side->stream, side->fork, 0));
    rc = kt_chain(x, x_bstride, b1, C, N, k, sqnorm, ws, idx_out, dist_out, nullptr, set_only, st);
    const int rc2 = kt_chain(x + (int64_t)b1 * x_bstride, x_bstride, b2, C, N, k, sqnorm + (int64_t)b1 * N, ws + kt_plan(b1, C, N).total,
                             idx_out + (int64_t)b1 * N * k, dist_out ? dist_out + (int64_t)b1 * N * k : nullptr, nullptr, set_only,
                             side->stream);
    // join even after a failed launch: a stream left forked would break an enclosing graph capture
    const cudaError_t e1 = cudaEventRecord(side->join, side->stream);
    const cudaError_t e2 = cudaStreamWaitEvent(st, side->join, 0);
    pdl_break();
    if (rc != GFS_OK) return rc;
    if (rc2 != GFS_OK) return rc2;
    GFS_CUDA_OK(e1);
    GFS_CUDA_OK(e2);
    return GFS_OK;
}

extern "C" int gfs_knn_tc_f32(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, void* workspace,
                              int64_t workspace_bytes, int32_t* idx_out, float* dist_out, void* stream) {
    return kt_run("gfs_knn_tc_f32", x, x_bstride, B, C, N, k, sqnorm, workspace, workspace_bytes, idx_out, dist_out, nullptr, false, stream);
}

extern "C" int gfs_knn_tc_set_f32(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, void* workspace,
                                  int64_t workspace_bytes, int32_t* idx_out, void* stream) {
    return kt_run("gfs_knn_tc_set_f32", x, x_bstride, B, C, N, k, sqnorm, workspace, workspace_bytes, idx_out, nullptr, nullptr, true, stream);
}

extern "C" int gfs_knn_tc_diag_f32(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, void* workspace,
                                   int64_t workspace_bytes, int32_t* idx_out, float* filter_out, int32_t* repair_flags_out,
                                   void* stream) {
    using namespace gfs;
    GFS_REQUIRE(filter_out && repair_flags_out, GFS_ERR_BAD_ARG, "gfs_knn_tc_diag_f32: null output");
    const int rc = kt_run("gfs_knn_tc_diag_f32", x, x_bstride, B, C, N, k, sqnorm, workspace, workspace_bytes, idx_out, nullptr, filter_out, false, stream);
    if (rc != GFS_OK) return rc;
    const KtPlan p = kt_plan(B, C, N);
    GFS_CUDA_OK(cudaMemcpyAsync(repair_flags_out, static_cast<uint8_t*>(workspace) + p.off_flags, (size_t)B * ((N + 63) / 64) * 4,
                                cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return GFS_OK;
}
