// fp32_tile.cuh -- the register-tiled fp32 FFMA core shared by kNN, the split first EdgeConv conv, the GW projection
// and the k-means E-step.  A CTA of 128 threads owns a 64 (rows) x 128 (cols) tile; thread (ty, tx) = (tid/16, tid%16)
// owns rows ty*8..ty*8+7 and columns {tx*4..tx*4+3} U {64+tx*4..64+tx*4+3}.  Operands live in shared memory
// channel-major (As[c][row], Bs[c][col]) so every dot product is ONE fma chain over c ascending -- the pinned
// evaluation order of oracle/gfs_oracle.c -- regardless of how the tile is scheduled.
#pragma once
#include "common.cuh"

namespace gfs {

constexpr int T_ROWS = 64;
constexpr int T_COLS = 128;
constexpr int T_THREADS = 128;

// acc[r][s] += sum_{c in [0,nc)} As[c][ty*8+r] * Bs[c][col(s)]
// Issued as packed FFMA2 (fma.rn.f32x2, new on sm_100): two IEEE fp32 fmas per instruction on adjacent columns, which
// halves the issue slots of the inner product (each element is still its own fma chain, so the pinned order holds).
__device__ __forceinline__ void tile_fma(const float* __restrict__ As, const float* __restrict__ Bs, int nc, int ty, int tx,
                                         float (&acc)[8][8]) {
    float2 acc2[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc2[r][s] = make_float2(acc[r][2 * s], acc[r][2 * s + 1]);
#pragma unroll 4
    for (int c = 0; c < nc; ++c) {
        const float4 a0 = *reinterpret_cast<const float4*>(As + c * T_ROWS + ty * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(As + c * T_ROWS + ty * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + c * T_COLS + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + c * T_COLS + 64 + tx * 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const float2 aa = make_float2(a[r], a[r]);
#pragma unroll
            for (int s = 0; s < 4; ++s) acc2[r][s] = __ffma2_rn(aa, b[s], acc2[r][s]);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            acc[r][2 * s] = acc2[r][s].x;
            acc[r][2 * s + 1] = acc2[r][s].y;
        }
}

// cp.async a (nrows x width) fp32 panel: dst[r][0..width) <- src[r*src_ld + col0 .. ), zero-filled past `limit` columns.
// width % 4 == 0; (src + r*src_ld + col0) must be 16-byte aligned.
__device__ __forceinline__ void load_panel_async(float* dst, int width, const float* src, int64_t src_ld, int nrows, int col0,
                                                 int limit, int tid) {
    const int cpr = width >> 2;   // 16-byte chunks per row
    for (int i = tid; i < nrows * cpr; i += T_THREADS) {
        const int r = i / cpr, q = i - r * cpr;
        const int col = col0 + q * 4;
        int bytes = (limit - col) * 4;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const float* s = src + (int64_t)r * src_ld + (bytes > 0 ? col : 0);
        cp_async16(dst + r * width + q * 4, s, bytes);
    }
}

// per-point squared norm as the same fma chain (c ascending) the tile core uses for a dot product: |x|^2 == dot(x, x)
static __global__ void sqnorm_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (n >= N) return;
    const float* p = x + (int64_t)b * bstride + n;
    float acc = 0.0f;
    for (int c = 0; c < C; ++c) {
        const float v = p[(int64_t)c * N];
        acc = fmaf(v, v, acc);
    }
    out[(int64_t)b * N + n] = acc;
}

}  // namespace gfs
