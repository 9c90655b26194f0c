// pointwise.cu -- per-point fp32 1x1 convolution on the shared FFMA core.
//
// Used for the algebraically split first EdgeConv conv (model/dgcnn.py:26-42 + :53):
//     W1 . cat(x_j - x_i, x_i) = Wa . x_j + (Wb - Wa) . x_i
// so one (C -> 128) per-point product gives P' = s1*(Wa x) and Q' = s1*((Wb-Wa) x) + t1 for every point once, instead
// of a (2C -> 64) product over every one of the N*k edges.  Output is point-major so that a neighbour's P' row is one
// contiguous 256-byte gather in the EdgeConv kernel.
#include "fp32_tile.cuh"

namespace gfs {

__global__ void __launch_bounds__(T_THREADS, 3)
pointwise_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, const float* __restrict__ wt,
                 const float* __restrict__ bias, int O, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);   // [C][64]
    float* Bs = As + C * T_ROWS;                      // [C][128]
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.y, n0 = blockIdx.x * T_ROWS, o0 = blockIdx.z * T_COLS;

    load_panel_async(As, T_ROWS, x + (int64_t)b * bstride, N, C, n0, N, tid);
    load_panel_async(Bs, T_COLS, wt, O, C, o0, O, tid);
    cp_async_commit();

    float acc[8][8];
    float bv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int v = 0; v < 4; ++v) bv[h * 4 + v] = bias ? bias[o0 + h * 64 + tx * 4 + v] : 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = bv[c];

    cp_async_wait<0>();
    __syncthreads();
    tile_fma(As, Bs, C, ty, tx, acc);

#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int n = n0 + ty * 8 + r;
        if (n >= N) continue;
        float* o = out + ((int64_t)b * N + n) * O + o0;
        *reinterpret_cast<float4*>(o + tx * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        *reinterpret_cast<float4*>(o + 64 + tx * 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
}

// The split first EdgeConv conv in the layout the EdgeConv kernel gathers from:
//   pb (B*N, 64) bf16 = P' - mu   and   q (B*N, 64) fp32 = Q' + mu,   mu = P'(shift of the block)
// Only the SUM P'[j] + Q'[i] is ever used, so any per-block constant mu may move from one to the other.  With mu = the
// image of a point inside the block's cloud (the mean of four spread-out points, as in knn_prep_kernel) the gathered term
// is a small difference: rounding it to bf16 costs 2^-9 of |P' - mu| instead of 2^-9 of |P'| -- measured on the bench
// model: no change of the layer's error against fp64 (it stays the bf16 rounding of conv2's operands) -- and a neighbour
// row shrinks from 256 to 128 bytes, which halves the L2 gather traffic that bounds the EdgeConv kernel.
__global__ void __launch_bounds__(T_THREADS, 3)
edge_pq_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, const float* __restrict__ wt,
               const float* __restrict__ bias, __nv_bfloat16* __restrict__ pb, float* __restrict__ qo) {
    pdl_enter();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);   // [C][64]
    float* Bs = As + C * T_ROWS;                      // [C][128]
    float* mux = Bs + C * T_COLS;                     // [C] the block's shift
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.y, n0 = blockIdx.x * T_ROWS;
    const float* xb = x + (int64_t)b * bstride;

    load_panel_async(As, T_ROWS, xb, N, C, n0, N, tid);
    load_panel_async(Bs, T_COLS, wt, T_COLS, C, 0, T_COLS, tid);
    cp_async_commit();
    if (tid < C) {
        const float* p = xb + (int64_t)tid * N;
        mux[tid] = 0.25f * ((__ldg(p) + __ldg(p + N / 4)) + (__ldg(p + N / 2) + __ldg(p + 3 * (N / 4))));
    }
    float acc[8][8];
    float bv[8];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int v = 0; v < 4; ++v) bv[h * 4 + v] = bias ? bias[h * 64 + tx * 4 + v] : 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = bv[c];

    cp_async_wait<0>();
    __syncthreads();
    tile_fma(As, Bs, C, ty, tx, acc);
    // mu for this thread's four P' columns
    float mu[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    for (int c = 0; c < C; ++c) {
        const float4 w = *reinterpret_cast<const float4*>(Bs + c * T_COLS + tx * 4);
        const float m = mux[c];
        mu[0] = fmaf(m, w.x, mu[0]);
        mu[1] = fmaf(m, w.y, mu[1]);
        mu[2] = fmaf(m, w.z, mu[2]);
        mu[3] = fmaf(m, w.w, mu[3]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int n = n0 + ty * 8 + r;
        if (n >= N) continue;
        const int64_t row = (int64_t)b * N + n;
        uint2 pk;
        pk.x = pack_bf16x2(acc[r][0] - mu[0], acc[r][1] - mu[1]);
        pk.y = pack_bf16x2(acc[r][2] - mu[2], acc[r][3] - mu[3]);
        *reinterpret_cast<uint2*>(pb + row * 64 + tx * 4) = pk;
        *reinterpret_cast<float4*>(qo + row * 64 + tx * 4) = make_float4(acc[r][4] + mu[0], acc[r][5] + mu[1], acc[r][6] + mu[2], acc[r][7] + mu[3]);
    }
}

}  // namespace gfs

extern "C" int gfs_edge_pq_f32(const float* x, int64_t x_bstride, int B, int C, int N, const float* wt, const float* bias,
                               void* pb, float* q, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x && wt && pb && q, GFS_ERR_BAD_ARG, "gfs_edge_pq_f32: null pointer");
    GFS_REQUIRE(B > 0 && C > 0 && N > 0, GFS_ERR_BAD_ARG, "gfs_edge_pq_f32: non-positive size");
    GFS_REQUIRE(C <= 64, GFS_ERR_UNSUPPORTED, "gfs_edge_pq_f32: C=%d > 64 is not built", C);
    GFS_REQUIRE(N % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wt) & 15) == 0 && (reinterpret_cast<uintptr_t>(pb) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(q) & 15) == 0,
                GFS_ERR_UNSUPPORTED, "gfs_edge_pq_f32: needs N %% 4 == 0 and 16-byte aligned pointers");
    const size_t smem = (size_t)C * (T_ROWS + T_COLS + 1) * sizeof(float);
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(edge_pq_kernel), 64 * (T_ROWS + T_COLS + 1) * sizeof(float)));
    launch_pdl(edge_pq_kernel, dim3((N + T_ROWS - 1) / T_ROWS, B), dim3(T_THREADS), smem, static_cast<cudaStream_t>(stream),
        x, x_bstride, C, N, wt, bias, static_cast<__nv_bfloat16*>(pb), q);
    GFS_LAUNCH_OK("edge_pq_kernel");
    return GFS_OK;
}

extern "C" int gfs_pointwise_f32(const float* x, int64_t x_bstride, int B, int C, int N, const float* wt, const float* bias,
                                 int O, float* out, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x && wt && out, GFS_ERR_BAD_ARG, "gfs_pointwise_f32: null pointer");
    GFS_REQUIRE(B > 0 && C > 0 && N > 0 && O > 0, GFS_ERR_BAD_ARG, "gfs_pointwise_f32: non-positive size");
    GFS_REQUIRE(C <= 64, GFS_ERR_UNSUPPORTED, "gfs_pointwise_f32: C=%d > 64 is not built", C);
    GFS_REQUIRE(O % T_COLS == 0, GFS_ERR_UNSUPPORTED, "gfs_pointwise_f32: O=%d must be a multiple of 128", O);
    GFS_REQUIRE(N % 4 == 0 && x_bstride % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wt) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                GFS_ERR_UNSUPPORTED, "gfs_pointwise_f32: needs N %% 4 == 0 and 16-byte aligned pointers");
    const size_t smem = (size_t)C * (T_ROWS + T_COLS) * sizeof(float);
    GFS_CUDA_OK(allow_smem(reinterpret_cast<const void*>(pointwise_kernel), 64 * (T_ROWS + T_COLS) * sizeof(float)));
    pointwise_kernel<<<dim3((N + T_ROWS - 1) / T_ROWS, B, O / T_COLS), T_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
        x, x_bstride, C, N, wt, bias, O, out);
    GFS_LAUNCH_OK("pointwise_kernel");
    return GFS_OK;
}
