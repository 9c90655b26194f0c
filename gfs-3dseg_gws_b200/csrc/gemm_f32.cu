// gemm_f32.cu -- generic fp32 GEMM on the packed-FFMA2 tile core, the workhorse of the TRAINING path
// (model/dgcnn.py and model/capl.py under model.train(): every Conv1d/Conv2d 1x1 forward, its data gradient and its
// weight gradient; training keeps fp32 end to end like the reference, tolerance 1e-3).
//
//   C[r, n] = bias[r] + sum_k Aop(k, r) * Bop(k, n)          r < R, n < Ncols, k < K, optionally batched
//     Aop(k, r) = a_trans ? A[r*lda + k] : A[k*lda + r]       (same for B with n)
//     C is stored row-major C[r*ldc + n] or, with c_trans, C[n*ldc + r]
//
// Activations live channel-major (C, M), so with  A = W^T  the forward conv, with  A = W  the data gradient and with
// both operands transposed the weight gradient (a contraction over the M points, split over CTAs and reduced in a fixed
// order) are all this one kernel.  64 x 128 output tile per CTA of 128 threads, K streamed through shared memory in
// chunks of 32.
#include "fp32_tile.cuh"

namespace gfs {

constexpr int GM_KC = 32;

// panel[kk][j] (j < width) <- op(k0 + kk, j0 + j), zero outside [0,K) x [0,lim)
__device__ __forceinline__ void gm_load_panel(float* panel, int width, const float* __restrict__ src, int64_t ld, int trans,
                                              int k0, int K, int j0, int lim, int tid) {
    if (!trans) {
        // source row k holds consecutive j: float4 along j when aligned, scalar otherwise
        const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((j0 & 3) == 0);
        const int cpr = width >> 2;
        for (int i = tid; i < GM_KC * cpr; i += T_THREADS) {
            const int kk = i / cpr, q = i - kk * cpr;
            const int k = k0 + kk, j = j0 + q * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K) {
                const float* p = src + (int64_t)k * ld + j;
                if (vec && j + 3 < lim) {
                    v = __ldg(reinterpret_cast<const float4*>(p));
                } else {
                    if (j < lim) v.x = __ldg(p);
                    if (j + 1 < lim) v.y = __ldg(p + 1);
                    if (j + 2 < lim) v.z = __ldg(p + 2);
                    if (j + 3 < lim) v.w = __ldg(p + 3);
                }
            }
            *reinterpret_cast<float4*>(panel + kk * width + q * 4) = v;
        }
    } else {
        // source row j holds consecutive k: each thread takes 4 consecutive k of one j (one float4 when aligned); lanes vary
        // along j -> conflict-free shared-memory stores
        const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((k0 & 3) == 0);
        for (int i = tid; i < (GM_KC / 4) * width; i += T_THREADS) {
            const int kq = i / width, jj = i - kq * width;
            const int j = j0 + jj, k = k0 + kq * 4;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (j < lim) {
                const float* p = src + (int64_t)j * ld + k;
                if (vec && k + 3 < K) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (k + e < K) v[e] = __ldg(p + e);
                }
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) panel[(kq * 4 + e) * width + jj] = v[e];
        }
    }
}

__global__ void __launch_bounds__(T_THREADS, 3)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, int a_trans, int64_t a_bs, const float* __restrict__ B, int64_t ldb,
                int b_trans, int64_t b_bs, float* __restrict__ C, int64_t ldc, int c_trans, int64_t c_bs,
                const float* __restrict__ bias, int R, int Ncols, int K, int splitk, int kper) {
    __shared__ __align__(16) float As[GM_KC * T_ROWS];
    __shared__ __align__(16) float Bs[GM_KC * T_COLS];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int n0 = blockIdx.x * T_COLS, r0 = blockIdx.y * T_ROWS;
    const int z = blockIdx.z;
    const int bz = z / splitk, sk = z - bz * splitk;
    A += (int64_t)bz * a_bs;
    B += (int64_t)bz * b_bs;
    const int kbeg = sk * kper;
    const int kend = (kbeg + kper) < K ? (kbeg + kper) : K;
    // split-K partials go to C viewed as [z][R][Ncols] (the caller passes a workspace and reduces afterwards)
    C += splitk > 1 ? (int64_t)z * R * Ncols : (int64_t)bz * c_bs;

    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int rr = r0 + ty * 8 + r;
        const float bv = (bias && splitk == 1 && rr < R) ? bias[rr] : 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = bv;
    }
    for (int k0 = kbeg; k0 < kend; k0 += GM_KC) {
        gm_load_panel(As, T_ROWS, A, lda, a_trans, k0, kend, r0, R, tid);
        gm_load_panel(Bs, T_COLS, B, ldb, b_trans, k0, kend, n0, Ncols, tid);
        __syncthreads();
        const int kn = (kend - k0) < GM_KC ? (kend - k0) : GM_KC;
        tile_fma(As, Bs, kn, ty, tx, acc);
        __syncthreads();
    }
    const bool row_major = splitk > 1 || !c_trans;
    const int64_t ld = splitk > 1 ? Ncols : ldc;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int rr = r0 + ty * 8 + r;
        if (rr >= R) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int nn = n0 + (c >> 2) * 64 + tx * 4 + (c & 3);
            if (nn >= Ncols) continue;
            if (row_major) C[(int64_t)rr * ld + nn] = acc[r][c];
            else C[(int64_t)nn * ld + rr] = acc[r][c];
        }
    }
}

__global__ void gemm_splitk_reduce_kernel(const float* __restrict__ part, int splitk, int R, int Ncols, const float* __restrict__ bias,
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

extern "C" int gfs_gemm_f32(const float* A, int64_t lda, int a_trans, int64_t a_bstride, const float* B, int64_t ldb, int b_trans,
                            int64_t b_bstride, float* C, int64_t ldc, int c_trans, int64_t c_bstride, const float* bias, int R,
                            int Ncols, int K, int batch, int splitk, float* workspace, int accumulate, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(A && B && C, GFS_ERR_BAD_ARG, "gfs_gemm_f32: null pointer");
    GFS_REQUIRE(R > 0 && Ncols > 0 && K > 0 && batch > 0 && splitk > 0, GFS_ERR_BAD_ARG, "gfs_gemm_f32: non-positive size");
    GFS_REQUIRE(splitk == 1 || (batch == 1 && workspace), GFS_ERR_BAD_ARG, "gfs_gemm_f32: split-K needs batch == 1 and a workspace");
    GFS_REQUIRE(splitk > 1 || !accumulate, GFS_ERR_UNSUPPORTED, "gfs_gemm_f32: accumulate is only built for the split-K reduction");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int kper = (K + splitk - 1) / splitk;
    kper = (kper + GM_KC - 1) / GM_KC * GM_KC;
    const dim3 grid((Ncols + T_COLS - 1) / T_COLS, (R + T_ROWS - 1) / T_ROWS, batch * splitk);
    GFS_REQUIRE(grid.y <= 65535 && grid.z <= 65535, GFS_ERR_UNSUPPORTED, "gfs_gemm_f32: grid too large");
    gemm_f32_kernel<<<grid, T_THREADS, 0, st>>>(A, lda, a_trans, a_bstride, B, ldb, b_trans, b_bstride, splitk > 1 ? workspace : C, ldc,
                                                c_trans, c_bstride, bias, R, Ncols, K, splitk, kper);
    GFS_LAUNCH_OK("gemm_f32_kernel");
    if (splitk > 1) {
        gemm_splitk_reduce_kernel<<<(R * Ncols + 255) / 256, 256, 0, st>>>(workspace, splitk, R, Ncols, bias, C, ldc, c_trans, accumulate);
        GFS_LAUNCH_OK("gemm_splitk_reduce_kernel");
    }
    return GFS_OK;
}
