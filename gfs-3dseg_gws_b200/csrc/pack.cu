// pack.cu -- layout conversion into the tcgen05-ready bf16 tile formats described in include/gfs3d.h.
#include "common.cuh"

namespace gfs {

// one thread per (k-block, row, 16-byte chunk): 8 consecutive K elements of one weight row
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ row_scale, int R, int K, int kblocks,
                                   uint8_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)kblocks * R * 8) return;
    const int q = (int)(i & 7);
    const int r = (int)((i >> 3) % R);
    const int kb = (int)((i >> 3) / R);
    const float sc = row_scale ? row_scale[r] : 1.0f;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int kcol = kb * 64 + q * 8 + e;
        v[e] = kcol < K ? w[(int64_t)r * K + kcol] * sc : 0.0f;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]);
    o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]);
    o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (int64_t)kb * R * 128 + sw128(r, q)) = o;
}

// fp32 channel-major -> bf16 act tiles.  CTA = (64-channel block, 128-point tile); reads are coalesced along points,
// the transpose goes through registers: thread t owns point row t and walks the 64 channels.
__global__ void __launch_bounds__(128)
cm_to_act_kernel(const float* __restrict__ x, int64_t bstride, int C, int N, int64_t M, uint8_t* __restrict__ act, int kblocks,
                 int kb0) {
    const int64_t m = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int cb = blockIdx.y;
    uint8_t* tile = act + ((int64_t)blockIdx.x * kblocks + kb0 + cb) * 16384;
    uint4 pk[8];
    if (m < M) {
        const int64_t b = m / N, n = m - b * N;
        const float* p = x + b * bstride + (int64_t)(cb * 64) * N + n;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = p[(int64_t)(q * 8 + e) * N];
            pk[q].x = pack_bf16x2(v[0], v[1]);
            pk[q].y = pack_bf16x2(v[2], v[3]);
            pk[q].z = pack_bf16x2(v[4], v[5]);
            pk[q].w = pack_bf16x2(v[6], v[7]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q) pk[q] = make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(tile + sw128(threadIdx.x, q)) = pk[q];
}

}  // namespace gfs

extern "C" int gfs_pack_weight_bf16(const float* w, const float* row_scale, int R, int K, void* out, void* stream) {
    using namespace gfs;
    GFS_REQUIRE(w && out, GFS_ERR_BAD_ARG, "gfs_pack_weight_bf16: null pointer");
    GFS_REQUIRE(R > 0 && K > 0 && R % 8 == 0, GFS_ERR_BAD_ARG, "gfs_pack_weight_bf16: R=%d must be a positive multiple of 8", R);
    GFS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, GFS_ERR_BAD_ARG, "gfs_pack_weight_bf16: out must be 16-byte aligned");
    const int kblocks = (K + 63) / 64;
    const int64_t total = (int64_t)kblocks * R * 8;
    pack_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, row_scale, R, K, kblocks, static_cast<uint8_t*>(out));
    GFS_LAUNCH_OK("pack_weight_kernel");
    return GFS_OK;
}

extern "C" int gfs_cm_to_act(const float* x, int64_t x_bstride, int B, int C, int N, void* act, int kblocks, int kb0,
                             void* stream) {
    using namespace gfs;
    GFS_REQUIRE(x && act, GFS_ERR_BAD_ARG, "gfs_cm_to_act: null pointer");
    GFS_REQUIRE(B > 0 && N > 0 && C > 0 && C % 64 == 0, GFS_ERR_BAD_ARG, "gfs_cm_to_act: C=%d must be a positive multiple of 64", C);
    GFS_REQUIRE(kb0 >= 0 && kb0 + C / 64 <= kblocks, GFS_ERR_BAD_ARG, "gfs_cm_to_act: column blocks out of range");
    const int64_t M = (int64_t)B * N;
    cm_to_act_kernel<<<dim3((unsigned)((M + 127) / 128), C / 64), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        x, x_bstride, C, N, M, static_cast<uint8_t*>(act), kblocks, kb0);
    GFS_LAUNCH_OK("cm_to_act_kernel");
    return GFS_OK;
}
