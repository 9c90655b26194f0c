// knn_tc.cu -- kNN graph with a tensor-core candidate filter and an exact fp32 finish (model/dgcnn.py:17-23).
//
// The pinned result (oracle/gfs_oracle.c: d = fmaf(2, dot, -xx_i) - xx_j with dot one fp32 fma chain, the k largest per row,
// ties -> ascending index) cannot be formed on tensor cores, but the tensor cores can tell cheaply which few candidates
// can possibly be among the k best:
//
//   1. knn_prep_kernel: distances do not change under a translation, so the filter works on coordinates shifted by a
//      point inside the block's cloud (the mean of four spread-out points; small norms = small absolute error).  Every
//      shifted coordinate is split into two bf16 terms x~ = h + l (+ r, |r| <= 2^-18 |x~|) and stored as [h | l] in the
//      UMMA K-major SWIZZLE_128B tile layout, next to -|x~_j|^2/2 + a_j and a tag (j << 16 | bf16(2 a_j)) per point and
//      a point-major fp32 copy of the ORIGINAL coordinates.
//   2. knn_tc_kernel: one CTA owns 256 query rows of one block (two UMMA M=128 tiles, operands resident in shared
//      memory) and streams the block's candidates in stages of 64 (TMA bulk copies, double buffered).  Per stage three
//      tcgen05.mma chains  h_i.h_j + h_i.l_j + l_i.h_j  leave  D' ~= x~_i.x~_j  in TMEM (fp32, four stages in flight).
//      Eight selection warps read TMEM with tcgen05.ld, ONE THREAD PER QUERY ROW (no shuffles, no cross-lane sorting):
//      every candidate whose filter value u = D' - |x~_j|^2/2 + a_j reaches the row's threshold is appended to the row's
//      cell array in shared memory (predicated 8-byte stores of {u, tag}; a few percent of the candidates).  When the
//      cells fill up, the new ones are folded into a k-deep sorted register list (min/max chains), the threshold becomes
//      (k-th best lower bound so far) - 2 a_i, and the cells below it are dropped.
//      Error bound (per PAIR, so that one far-away point does not loosen the filter for every row):
//        |v(i,j) - D(i,j)| <= a_i + a_j,   a = 2^-15 |x~|^2 + (C+4) 2^-25 |x|^2   per point
//      (bf16 split residual <= 3 * 2^-18 |x~_i||x~_j|, <= 192 fp32 accumulations in the tensor core at <= 1 ulp each,
//      |x~_i||x~_j| <= (|x~_i|^2+|x~_j|^2)/2, plus the rounding of the pinned fp32 chain on the original coordinates;
//      tests/test_gpu_knn_tc.py measures the observed error against it).  The filter value carries +a_j, i.e. it is an
//      UPPER bound u of D up to the row constant a_i; each cell also carries 2 a_j (bf16, rounded up) next to the column
//      index, so the sorted list ranks the LOWER bounds w = u - 2 a_j.  The k-th largest lower bound minus 2 a_i cannot
//      exceed the exact k-th best D, hence a candidate with u below it cannot be among the exact top k.
//   3. knn_finish_kernel: the survivors (k plus a handful) get their EXACT pinned distance from the point-major fp32
//      copy, one warp per row and one lane per survivor (32 per round), are ranked by (d desc, index asc) and written out:
//      bit-identical to knn.cu.
//   Rows with more than P-8 candidates inside the bound (duplicated points; feature spaces so collapsed that the pinned
//   chain's own rounding is of the order of the neighbour spacing) move their cells' columns to the row's survivor list in
//   global memory and carry on; only a row that collects more than 256 survivors (a flood of exact ties) raises a flag
//   for its 64-row tile, and knn.cu's exact kernel redoes just those tiles inside the same call.
#include "common.cuh"

namespace gfs {

constexpr int KT_ROWS = 256;          // query rows per CTA
constexpr int KT_COLS = 64;           // candidates per stage
constexpr int KT_TST = 4;             // TMEM stages: 2 row tiles x 4 stages x 64 fp32 columns = 512 columns
constexpr int KT_THREADS = 384;       // warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4-11 selection
constexpr int KT_P = 63;              // candidate cells (8 bytes) per row (63: the control block has to fit next to them)
constexpr int KT_SURV = 256;          // survivors per row the finish kernel can take (32 per round, one per lane)

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
// One CTA per 128 points.  The (C x 128) slab is staged in shared memory with coalesced loads; thread t then owns point t
// for the two norm chains, and the operand tiles / the point-major copy are written with (row, 16-byte chunk) work items
// so that consecutive threads write consecutive bytes.
constexpr int KP_LD = 129;   // padded row of the staged slab
__global__ void __launch_bounds__(128)
knn_prep_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, int Npad, int Cp16, int KB, int CPT,
                uint8_t* __restrict__ ops, float* __restrict__ xp, float* __restrict__ nh, uint32_t* __restrict__ tag,
                float* __restrict__ sqnorm) {
    __shared__ float xs[64 * KP_LD];
    __shared__ float mus[64];
    const int t = threadIdx.x, rt = blockIdx.x, b = blockIdx.y;
    const int n0 = rt * 128, n = n0 + t;
    const bool valid = n < N;
    const float* xb = x + (int64_t)b * bstride;
    // the shift: any vector is valid (distances are translation invariant), it only has to be the SAME for every point of
    // the block and close to the data.  The mean of four spread-out points costs four loads and no extra kernel.
    if (t < 64) {
        const float* p = xb + (int64_t)t * N;
        mus[t] = t < C ? 0.25f * ((__ldg(p) + __ldg(p + N / 4)) + (__ldg(p + N / 2) + __ldg(p + 3 * (N / 4)))) : 0.0f;
    }
    for (int c = 0; c < Cp16; ++c) xs[c * KP_LD + t] = (valid && c < C) ? __ldg(xb + (int64_t)c * N + n) : 0.0f;
    __syncthreads();

    float xx = 0.0f, cc = 0.0f;
    for (int c = 0; c < C; ++c) {
        const float v = xs[c * KP_LD + t];
        const float u = v - mus[c];
        xx = fmaf(v, v, xx);                 // same chain as sqnorm_kernel
        cc = fmaf(u, u, cc);
    }
    if (!valid) cc = 0.0f;                   // padding rows: all-zero operands (their v - mu would not be zero)
    // a = the point's share of the pair error bound  |v(i,j) - D(i,j)| <= a_i + a_j  (see the header):
    //   2^-15 |x~|^2  covers the bf16 split residual and the tensor core's fp32 accumulation (with |x~_i||x~_j| <= (|x~_i|^2+|x~_j|^2)/2),
    //   (C+4) 2^-25 |x|^2  the rounding of the pinned fp32 chain on the original coordinates.
    // The filter value carries +a_j (an upper bound of D), the tag carries 2 a_j rounded UP to bf16 (to get the lower bound back).
    const float a = 0x1p-15f * cc + (float)(C + 4) * 0x1p-25f * xx;
    const __nv_bfloat16 dl = __float2bfloat16_ru(2.0f * a);
    nh[(int64_t)b * Npad + n] = valid ? -0.5f * cc + a : -INFINITY;
    tag[(int64_t)b * Npad + n] = ((uint32_t)n << 16) | (uint32_t)__bfloat16_as_ushort(dl);
    if (valid) sqnorm[(int64_t)b * N + n] = xx;

    // operand tiles: columns [h: 0..Cp16) [l: Cp16..2 Cp16), 8 columns per 16-byte chunk
    uint8_t* tiles = ops + ((int64_t)b * (Npad / 128) + rt) * KB * 16384;
    const int cpr = Cp16 >> 3;               // chunks per row and per part
    for (int i = t; i < 128 * cpr; i += 128) {
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
        *reinterpret_cast<uint4*>(tiles + (ch >> 6) * 16384 + sw128(r, (ch & 63) >> 3)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        *reinterpret_cast<uint4*>(tiles + (cl >> 6) * 16384 + sw128(r, (cl & 63) >> 3)) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
    }
    // point-major fp32 copy of the ORIGINAL coordinates, rows of CPT floats (zero padded)
    float* xrows = xp + ((int64_t)b * Npad + n0) * CPT;
    const int fpr = CPT >> 2;
    for (int i = t; i < 128 * fpr; i += 128) {
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
    float nh[KT_TST][KT_COLS];        // -|x~_j|^2/2 + a_j of the stage's candidates
    uint32_t tag[KT_TST][KT_COLS];    // (j << 16) | bf16(2 a_j)
    uint64_t a_full, b_full[2], b_empty[2], d_full[KT_TST], d_empty[KT_TST];
    uint32_t tmem_base;
};

template <int KL>
__device__ __forceinline__ void kt_insert(float (&L)[KL], float v) {
#pragma unroll
    for (int i = 0; i < KL; ++i) {
        const float hi = fmaxf(L[i], v);
        v = fminf(L[i], v);
        L[i] = hi;
    }
}
template <int KL>
__device__ __forceinline__ float kt_kth(const float (&L)[KL], int k) {   // L[k-1] without indexing registers dynamically
    float r = L[0];
#pragma unroll
    for (int i = 1; i < KL; ++i) r = (i == k - 1) ? L[i] : r;
    return r;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];\n" : "=r"(r.x), "=r"(r.y) : "r"(a));
    return r;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};\n" ::"r"(a), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ u64 lds_u64(uint32_t a) {
    u64 r;
    asm volatile("ld.shared.u64 %0, [%1];\n" : "=l"(r) : "r"(a));
    return r;
}

// One selection thread's cells: cell e of row r is the 8 bytes at cells + (e * 256 + r) * 8 = {filter value bits, column}.
struct KtCells {
    uint32_t c0;      // address of the row's cell 0
    uint32_t end;     // next free cell
    uint32_t mark;    // cells [c0, mark) have already been folded into the sorted list
    uint16_t* spill;  // the row's survivor list in global memory: cells that had to make room go there (column only)
    int nspill;
};
constexpr uint32_t KT_STEP = KT_ROWS * 8;

// Bring the row's sorted list L up to date with the cells appended since the last call, take the k-th best as the new
// bound, drop every cell below (bound - margin).  Returns the new threshold.
template <int KL>
__device__ __forceinline__ float kt_refresh(float (&L)[KL], KtCells& s, float margin, int k) {
    for (uint32_t a = s.mark; a < s.end; a += 2 * KT_STEP) {
        // two values per trip: the two dependent min/max chains interleave
        // cell = {u = upper bound of D(i,j) up to a row constant, (j << 16) | bf16(2 a_j)}; the list ranks the LOWER bounds
        const uint2 e0 = lds64(a);
        const float v0 = __uint_as_float(e0.x) - __uint_as_float(e0.y << 16);
        float v1 = -INFINITY;
        if (a + KT_STEP < s.end) {
            const uint2 e1 = lds64(a + KT_STEP);
            v1 = __uint_as_float(e1.x) - __uint_as_float(e1.y << 16);
        }
        if (fmaxf(v0, v1) > L[KL - 1]) {
            kt_insert<KL>(L, v0);
            kt_insert<KL>(L, v1);
        }
    }
    const float thr = fmaxf(kt_kth<KL>(L, k) - margin, -3.402823466e38f);
    uint32_t w = s.c0;
    for (uint32_t a = s.c0; a < s.end; a += KT_STEP) {
        const uint2 e = lds64(a);
        if (__uint_as_float(e.x) >= thr) {
            sts64(w, e.x, e.y);
            w += KT_STEP;
        }
    }
    s.end = s.mark = w;
    return thr;
}

// Filter one chunk of 32 candidates (v = filter values, columns jb..jb+31) into the row's cells.
template <int KL>
__device__ __forceinline__ void kt_filter32(const float (&v)[32], const uint32_t* __restrict__ tags, float (&L)[KL], KtCells& cs,
                                            float& thr, bool& ovf, float margin, int k) {
    const uint32_t trigger = cs.c0 + (KT_P - 8) * KT_STEP;   // end > trigger  <=>  fewer than 8 free cells
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        if (__any_sync(0xffffffffu, cs.end > trigger)) {
            const float t = kt_refresh<KL>(L, cs, margin, k);
            if (cs.end > trigger) {
                // Still no room: more than P-8 candidates inside the error bound.  They are survivors whatever comes next
                // (the bound only rises, but their filter values are not kept): move their columns to the row's list in
                // global memory and go on with empty cells; the sorted list keeps their lower bounds.  A row that collects
                // more than KT_SURV this way (a flood of exact ties) is left to the exact repair pass.
                const int nc = (int)((cs.end - cs.c0) / KT_STEP);
                if (!ovf && cs.nspill + nc + KT_P <= KT_SURV) {
                    for (int e = 0; e < nc; ++e) cs.spill[cs.nspill + e] = (uint16_t)(lds64(cs.c0 + e * KT_STEP).y >> 16);
                    cs.nspill += nc;
                } else {
                    ovf = true;
                }
                cs.end = cs.mark = cs.c0;
            }
            thr = ovf ? INFINITY : t;
        }
        const uint4 t0 = *reinterpret_cast<const uint4*>(tags + g * 8), t1 = *reinterpret_cast<const uint4*>(tags + g * 8 + 4);
        const uint32_t tg[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (v[g * 8 + c] >= thr) {
                sts64(cs.end, __float_as_uint(v[g * 8 + c]), tg[c]);
                cs.end += KT_STEP;
            }
        }
    }
}

template <int KL>
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc_kernel(const uint8_t* __restrict__ ops, const float* __restrict__ nh, const uint32_t* __restrict__ tag,
              int* __restrict__ flags, int N, int Npad, int Cp16, int KB, int k,
              uint16_t* __restrict__ surv, int* __restrict__ surv_cnt, float* __restrict__ dbg) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    // [A: 2 row tiles x KB k-blocks x 16 KiB][B: 2 stages x KB x 8 KiB (64 candidate rows)][cells: P x 256 x 8 B][ctl]
    unsigned char* sA = base;
    const uint32_t a_bytes = (uint32_t)KB * 16384u;   // one row tile of the query operand
    const uint32_t b_bytes = (uint32_t)KB * 8192u;    // one candidate stage
    unsigned char* sB0 = sA + 2 * a_bytes;
    unsigned char* cells = sB0 + 2 * b_bytes;
    KtCtl& s = *reinterpret_cast<KtCtl*>(cells + KT_P * KT_ROWS * 8);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, q0 = blockIdx.x * KT_ROWS;
    const int nst = Npad / KT_COLS;
    const uint8_t* blk = ops + (int64_t)b * (Npad / 128) * a_bytes;

    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&s.a_full, 1);
            for (int i = 0; i < 2; ++i) {
                mbar_init(&s.b_full[i], 1);
                mbar_init(&s.b_empty[i], 1);
            }
            for (int i = 0; i < KT_TST; ++i) {
                mbar_init(&s.d_full[i], 1);
                mbar_init(&s.d_empty[i], 8);
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
    // the eight selection warps carry the register-resident lists and a prefetched TMEM chunk: (setmaxnreg: 4 warps x 40 + 8 warps x 224 registers per thread)
    // (setmaxnreg at the top of each role branch below)

    if (warp == 0) {
        // =============================== TMA producer ===============================
        reg_dealloc<40>();
        if (lane == 0) {
            mbar_arrive_expect_tx(&s.a_full, 2u * a_bytes);
            tma_load_1d(sA, blk + (int64_t)(q0 / 128) * a_bytes, 2u * a_bytes, &s.a_full);
            for (int st = 0; st < nst; ++st) {
                const int buf = st & 1, ts = st % KT_TST;
                mbar_wait(&s.b_empty[buf], ((st >> 1) & 1) ^ 1);
                mbar_wait(&s.d_empty[ts], ((st / KT_TST) & 1) ^ 1);   // the -|x_j|^2/2 slot is read by the selection warps
                mbar_arrive_expect_tx(&s.b_full[buf], b_bytes + KT_COLS * 8u);
                // candidates [st*64, st*64+64): half of row tile st/2, every k-block
                const uint8_t* src = blk + (int64_t)(st >> 1) * a_bytes + (st & 1) * 8192;
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_1d(sB0 + buf * b_bytes + kb * 8192, src + (int64_t)kb * 16384, 8192u, &s.b_full[buf]);
                tma_load_1d(s.nh[ts], nh + (int64_t)b * Npad + st * KT_COLS, KT_COLS * 4u, &s.b_full[buf]);
                tma_load_1d(s.tag[ts], tag + (int64_t)b * Npad + st * KT_COLS, KT_COLS * 4u, &s.b_full[buf]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        reg_dealloc<40>();
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, KT_COLS);
            const int ksteps = Cp16 >> 4;
            mbar_wait(&s.a_full, 0);
            for (int st = 0; st < nst; ++st) {
                const int buf = st & 1, ts = st % KT_TST;
                mbar_wait(&s.d_empty[ts], ((st / KT_TST) & 1) ^ 1);
                mbar_wait(&s.b_full[buf], (st >> 1) & 1);
                tc_fence_after();
                const uint32_t bB = smem_u32(sB0 + buf * b_bytes);
#pragma unroll 1
                for (int mt = 0; mt < 2; ++mt) {
                    const uint32_t aB = smem_u32(sA + mt * a_bytes);
                    const uint32_t d = tmem + mt * 256 + ts * KT_COLS;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const int ch = ks * 16, cl = Cp16 + ks * 16;
                        const uint32_t ih = (uint32_t)(ch & 63) * 2u, il = (uint32_t)(cl & 63) * 2u;
                        const uint64_t ah = umma_desc_sw128(aB + (uint32_t)(ch >> 6) * 16384u + ih);
                        const uint64_t al = umma_desc_sw128(aB + (uint32_t)(cl >> 6) * 16384u + il);
                        const uint64_t bh = umma_desc_sw128(bB + (uint32_t)(ch >> 6) * 8192u + ih);
                        const uint64_t bl = umma_desc_sw128(bB + (uint32_t)(cl >> 6) * 8192u + il);
                        umma_bf16(d, ah, bh, idesc, ks ? 1u : 0u);
                        umma_bf16(d, ah, bl, idesc, 1u);
                        umma_bf16(d, al, bh, idesc, 1u);
                    }
                }
                umma_commit(&s.b_empty[buf]);
                umma_commit(&s.d_full[ts]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // =============================== selection: one thread per query row ===============================
        reg_alloc<224>();
        const int mt = (warp - 4) >> 2, quarter = warp & 3;
        const int row = mt * 128 + quarter * 32 + lane;
        const int n = q0 + row;
        const bool valid = n < N;
        const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + mt * 256;
        KtCells cs;
        cs.c0 = smem_u32(cells) + row * 8;
        cs.end = cs.mark = cs.c0;
        cs.spill = surv + ((int64_t)b * N + (valid ? n : 0)) * KT_SURV;
        cs.nspill = 0;
        // margin = 2 a_i (the row's own share of the pair error bound, taken from its tag: rounded up)
        const float margin = valid ? __uint_as_float(tag[(int64_t)b * Npad + n] << 16) : 0.0f;
        bool ovf = !valid;          // "this row takes no more candidates": padding rows, and rows that overflowed
        float L[KL];
#pragma unroll
        for (int i = 0; i < KL; ++i) L[i] = -INFINITY;
        // -FLT_MAX, not -inf: padded candidates carry -inf and must never pass, not even while the bound is unknown
        float thr = valid ? -3.402823466e38f : INFINITY;

        // stages of 64 candidates = two chunks of 32; the TMEM load of the next chunk is in flight while one is filtered
        uint32_t r[32];
        float v[32];
        mbar_wait(&s.d_full[0], 0);
        tc_fence_after();
        tmem_ld32(tbase, r);
        for (int st = 0; st < nst; ++st) {
            const int ts = st % KT_TST;
            const float* nhs = s.nh[ts];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld_wait32(r);
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 h4 = *reinterpret_cast<const float4*>(nhs + half * 32 + c);
                    // packed adds (add.f32x2): two IEEE fp32 additions per issue slot
                    const float2 s0 = __fadd2_rn(make_float2(__uint_as_float(r[c]), __uint_as_float(r[c + 1])), make_float2(h4.x, h4.y));
                    const float2 s1 = __fadd2_rn(make_float2(__uint_as_float(r[c + 2]), __uint_as_float(r[c + 3])), make_float2(h4.z, h4.w));
                    v[c] = s0.x;
                    v[c + 1] = s0.y;
                    v[c + 2] = s1.x;
                    v[c + 3] = s1.y;
                }
                if (half == 0) {
                    tmem_ld32(tbase + ts * KT_COLS + 32, r);
                } else if (st + 1 < nst) {
                    const int ts1 = (st + 1) % KT_TST;
                    mbar_wait(&s.d_full[ts1], ((st + 1) / KT_TST) & 1);
                    tc_fence_after();
                    tmem_ld32(tbase + ts1 * KT_COLS, r);
                }
                const uint32_t jb = (uint32_t)(st * KT_COLS + half * 32);
                if (dbg && valid) {   // diagnostic entry only: dump the filter value v = D' - |x~_j|^2/2
                    float* o = dbg + ((int64_t)b * N + n) * Npad + jb;
#pragma unroll
                    for (int c = 0; c < 32; ++c) o[c] = v[c];
                }
                kt_filter32<KL>(v, s.tag[ts] + half * 32, L, cs, thr, ovf, margin, k);
            }
            // hand the TMEM stage (and its -|x~_j|^2/2 slot) back.  Deliberately at the END of the stage: with four stages in
            // flight nothing waits for it, and the values of the second half are certainly in registers by now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s.d_empty[ts]);
        }

        // ---- hand-off: the cells that survive the exact k-th best filter value go to knn_finish_kernel ----
        kt_refresh<KL>(L, cs, margin, k);   // cells = {v >= k-th best - margin}
        if (valid) {
            const int ns = (int)((cs.end - cs.c0) / KT_STEP);
            const int64_t g = (int64_t)b * N + n;
            if (ovf) {
                surv_cnt[g] = 0;
                flags[(int64_t)b * ((N + 63) / 64) + n / 64] = 1;
            } else {
                surv_cnt[g] = cs.nspill + ns;     // <= KT_SURV: a spill is only taken while nspill + cells + P fits
                uint16_t* o = cs.spill + cs.nspill;
                for (int e = 0; e < ns; ++e) o[e] = (uint16_t)(lds64(cs.c0 + e * KT_STEP).y >> 16);
            }
        }
    } else {
        reg_dealloc<40>();   // warps 2-3 only complete the first warpgroup
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// finish: exact pinned distances of the survivors, sorted, written out.  One warp per query row, one lane per survivor.
// The survivors' fp32 rows are gathered with coalesced loads (2 or 8 rows per instruction) into a padded staging tile,
// each lane then runs the pinned fma chain over its own row; a warp-wide bitonic sort of 64-bit keys
// (orderable d << 32 | ~j) puts the k best first: nearest first, ties -> ascending index.
// ---------------------------------------------------------------------------------------------------------------
template <int CPT, int KF_WARPS>
__global__ void __launch_bounds__(KF_WARPS * 32)
knn_finish_kernel(const float* __restrict__ xp, const float* __restrict__ sqnorm, const uint16_t* __restrict__ surv,
                  const int* __restrict__ surv_cnt, int N, int Npad, int k, int64_t rows, int32_t* __restrict__ idx_out,
                  float* __restrict__ dist_out) {
    constexpr int RS = CPT + 4;            // padded staging row: conflict-free LDS.128 across lanes
    constexpr int LPR = CPT / 4;           // lanes that fetch one row
    constexpr int RPI = 32 / LPR;          // rows per load instruction
    __shared__ __align__(16) float stg_all[KF_WARPS][32 * RS];
    __shared__ __align__(16) unsigned long long kbuf[KF_WARPS][KT_SURV];
    __shared__ __align__(16) float xi_all[KF_WARPS][CPT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* stg = stg_all[warp];
    const int sub = lane % LPR, grp = lane / LPR;
    // software pipeline over the warp's rows: the next row's survivor list is in flight while this row is finished
    const int64_t gstep = (int64_t)gridDim.x * KF_WARPS;
    int64_t g = (int64_t)blockIdx.x * KF_WARPS + warp;
    int ns_n = 0;
    uint32_t jl_n = 0, jh_n = 0;
    if (g < rows) {
        ns_n = __ldg(surv_cnt + g);
        jl_n = __ldg(surv + g * KT_SURV + lane);
        jh_n = __ldg(surv + g * KT_SURV + 32 + lane);
    }
    for (; g < rows; g += gstep) {
        const int ns = ns_n;
        const uint32_t jl = jl_n, jh = jh_n;
        if (g + gstep < rows) {
            ns_n = __ldg(surv_cnt + g + gstep);
            jl_n = __ldg(surv + (g + gstep) * KT_SURV + lane);
            jh_n = __ldg(surv + (g + gstep) * KT_SURV + 32 + lane);
        }
        if (ns == 0) continue;                       // flagged for the exact repair pass
        const int64_t b = g / N;
        const float* xpb = xp + b * Npad * CPT;
        // the query row itself goes to shared memory (broadcast reads in the chain): holding it in 64 registers per lane
        // would halve the number of resident warps, and this kernel lives on latency hiding
        float* xis = xi_all[warp];
        if (lane < LPR) *reinterpret_cast<float4*>(xis + lane * 4) = __ldg(reinterpret_cast<const float4*>(xpb + (g - b * N) * CPT + lane * 4));
        const float xxi = __ldg(sqnorm + g);
        u64* keys = reinterpret_cast<u64*>(kbuf[warp]);
        for (int half = 0; half * 32 < ns; ++half) {   // further rounds only for rows with more than 32 survivors
            const int e = half * 32 + lane;
            uint32_t j = 0u;                             // entries past ns are uninitialised memory
            if (e < ns) j = half == 0 ? jl : half == 1 ? jh : (uint32_t)__ldg(surv + g * KT_SURV + e);
            float4 t[32 / RPI];
#pragma unroll
            for (int it = 0; it < 32 / RPI; ++it) {
                const uint32_t jj = __shfl_sync(0xffffffffu, j, it * RPI + grp);
                t[it] = __ldg(reinterpret_cast<const float4*>(xpb + (int64_t)jj * CPT + sub * 4));
            }
            const float xxj = __ldg(sqnorm + b * N + j);
            __syncwarp();                            // the previous chains are done with the staging tile
#pragma unroll
            for (int it = 0; it < 32 / RPI; ++it) *reinterpret_cast<float4*>(stg + (it * RPI + grp) * RS + sub * 4) = t[it];
            __syncwarp();
            float dot = 0.0f;
#pragma unroll
            for (int c = 0; c < CPT; c += 4) {
                const float4 q = *reinterpret_cast<const float4*>(stg + lane * RS + c);
                const float4 a = *reinterpret_cast<const float4*>(xis + c);
                dot = fmaf(a.x, q.x, dot);
                dot = fmaf(a.y, q.y, dot);
                dot = fmaf(a.z, q.z, dot);
                dot = fmaf(a.w, q.w, dot);
            }
            const float d = fmaf(2.0f, dot, -xxi) - xxj;
            // 64-bit key: larger = nearer, equal distances -> smaller index first; 0 = empty slot (below every real key)
            keys[e] = e < ns ? (((u64)kt_ord_key(d) << 32) | (u64)(~j)) : 0ull;
        }
        __syncwarp();
        // rank by counting (independent broadcast reads: no dependent shuffle network), keys are all distinct
        const int nk = (ns + 3) & ~3;                  // slots up to the next multiple of 32 hold 0 = never greater
        for (int half = 0; half * 32 < ns; ++half) {
            const u64 mine = keys[half * 32 + lane];
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
            if (mine != 0ull && rank < k) {
                idx_out[g * k + rank] = (int32_t)(~(uint32_t)mine);
                if (dist_out) dist_out[g * k + rank] = kt_ord_val((uint32_t)(mine >> 32));
            }
        }
        __syncwarp();                                // keys are reused by the next row
    }
}

struct KtPlan {
    int Npad, Cp16, KB, CPT;
    size_t off_xp, off_nh, off_tag, off_surv, off_cnt, off_flags, total, zero_bytes;
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
    p.off_tag = o;
    o += (size_t)B * p.Npad * 4;
    p.off_surv = o;
    o += (size_t)B * N * KT_SURV * 2;
    p.off_cnt = o;
    o += (size_t)B * N * 4;
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

static int kt_launch(const KtPlan& p, uint8_t* ws, const float* sqnorm, int B, int C, int N, int k, int32_t* idx_out, float* dist_out,
                     float* dbg, cudaStream_t st) {
    const size_t smem = (size_t)2 * p.KB * 16384 + (size_t)2 * p.KB * 8192 + (size_t)KT_P * KT_ROWS * 8 + sizeof(KtCtl) + 1024;
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(knn_tc_kernel<20>), smem));
    uint16_t* surv = reinterpret_cast<uint16_t*>(ws + p.off_surv);
    int* cnt = reinterpret_cast<int*>(ws + p.off_cnt);
    knn_tc_kernel<20><<<dim3(p.Npad / KT_ROWS, B), KT_THREADS, smem, st>>>(
        ws, reinterpret_cast<const float*>(ws + p.off_nh), reinterpret_cast<const uint32_t*>(ws + p.off_tag),
        reinterpret_cast<int*>(ws + p.off_flags), N, p.Npad, p.Cp16, p.KB, k, surv, cnt, dbg);
    GFS_LAUNCH_OK("knn_tc_kernel");
    const int64_t rows = (int64_t)B * N;
    const int sms = sm_count();
    GFS_REQUIRE(sms > 0, GFS_ERR_CUDA, "gfs_knn_tc_f32: cannot query the device");
    const float* xp = reinterpret_cast<const float*>(ws + p.off_xp);
    if (p.CPT == 16) {
        const int64_t want = (rows + 7) / 8;
        const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
        knn_finish_kernel<16, 8><<<grid, 256, 0, st>>>(xp, sqnorm, surv, cnt, N, p.Npad, k, rows, idx_out, dist_out);
    } else {
        const int64_t want = (rows + 3) / 4;
        const int grid = (int)(want < (int64_t)sms * 16 ? want : (int64_t)sms * 16);
        knn_finish_kernel<64, 4><<<grid, 128, 0, st>>>(xp, sqnorm, surv, cnt, N, p.Npad, k, rows, idx_out, dist_out);
    }
    GFS_LAUNCH_OK("knn_finish_kernel");
    return GFS_OK;
}

}  // namespace gfs

extern "C" int64_t gfs_knn_tc_workspace_bytes(int B, int C, int N) {
    if (B <= 0 || C <= 0 || N <= 0 || C > 64) return 0;
    return (int64_t)gfs::kt_plan(B, C, N).total;
}

static int kt_run(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, void* workspace,
                  int64_t workspace_bytes, int32_t* idx_out, float* dist_out, float* dbg, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x && sqnorm && idx_out && workspace, GFS_ERR_BAD_ARG, "gfs_knn_tc_f32: null pointer");
    GFS_REQUIRE(B > 0 && C > 0 && N > 0 && k > 0, GFS_ERR_BAD_ARG, "gfs_knn_tc_f32: non-positive size (B=%d C=%d N=%d k=%d)", B, C, N, k);
    GFS_REQUIRE(k <= N, GFS_ERR_BAD_ARG, "gfs_knn_tc_f32: k=%d exceeds N=%d", k, N);
    GFS_REQUIRE(k <= 20, GFS_ERR_UNSUPPORTED, "gfs_knn_tc_f32: k=%d > 20 is not built (use gfs_knn_f32)", k);
    GFS_REQUIRE(C <= 64, GFS_ERR_UNSUPPORTED, "gfs_knn_tc_f32: C=%d > 64 is not built", C);
    GFS_REQUIRE(N <= 65535, GFS_ERR_UNSUPPORTED, "gfs_knn_tc_f32: N=%d > 65535 (16-bit candidate slots)", N);
    GFS_REQUIRE(N % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, GFS_ERR_UNSUPPORTED,
                "gfs_knn_tc_f32: needs N %% 4 == 0 and 16-byte aligned rows (N=%d)", N);
    const KtPlan p = kt_plan(B, C, N);
    GFS_REQUIRE(workspace_bytes >= (int64_t)p.total && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0, GFS_ERR_BAD_ARG,
                "gfs_knn_tc_f32: workspace of %lld bytes (256-byte aligned) needed, got %lld", (long long)p.total,
                (long long)workspace_bytes);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    GFS_CUDA_OK(cudaMemsetAsync(ws + p.off_flags, 0, p.zero_bytes, st));
    knn_prep_kernel<<<dim3(p.Npad / 128, B), 128, 0, st>>>(x, x_bstride, C, N, p.Npad, p.Cp16, p.KB, p.CPT, ws,
                                                           reinterpret_cast<float*>(ws + p.off_xp),
                                                           reinterpret_cast<float*>(ws + p.off_nh),
                                                           reinterpret_cast<uint32_t*>(ws + p.off_tag), sqnorm);
    GFS_LAUNCH_OK("knn_prep_kernel");
    const int rc = kt_launch(p, ws, sqnorm, B, C, N, k, idx_out, dist_out, dbg, st);
    if (rc != GFS_OK) return rc;
    return knn_exact_flagged(x, x_bstride, B, C, N, k, sqnorm, reinterpret_cast<const int*>(ws + p.off_flags), idx_out, dist_out, st);
}

extern "C" int gfs_knn_tc_f32(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, void* workspace,
                              int64_t workspace_bytes, int32_t* idx_out, float* dist_out, void* stream) {
    return kt_run(x, x_bstride, B, C, N, k, sqnorm, workspace, workspace_bytes, idx_out, dist_out, nullptr, stream);
}

extern "C" int gfs_knn_tc_diag_f32(const float* x, int64_t x_bstride, int B, int C, int N, int k, float* sqnorm, void* workspace,
                                   int64_t workspace_bytes, int32_t* idx_out, float* filter_out, int32_t* repair_flags_out,
                                   void* stream) {
    using namespace gfs;
    GFS_REQUIRE(filter_out && repair_flags_out, GFS_ERR_BAD_ARG, "gfs_knn_tc_diag_f32: null output");
    const int rc = kt_run(x, x_bstride, B, C, N, k, sqnorm, workspace, workspace_bytes, idx_out, nullptr, filter_out, stream);
    if (rc != GFS_OK) return rc;
    const KtPlan p = kt_plan(B, C, N);
    GFS_CUDA_OK(cudaMemcpyAsync(repair_flags_out, static_cast<uint8_t*>(workspace) + p.off_flags, (size_t)B * ((N + 63) / 64) * 4,
                                cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return GFS_OK;
}
