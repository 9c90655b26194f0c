// hist.cu -- joint histogram of two integer label streams, the reduction behind two callers of the hot path:
//
//   runs/eval.py:31-48   evaluate_metric_GFS: per-point Python loop over (gt, pred) -> the confusion matrix J[gt][pred]
//                        (gt_classes = row sums, positive_classes = column sums, true positives = diagonal)
//   train.py:156-218     collect_base_class_gp_coding_sum: per block, per class `sum(one_hot_gw[:, label == cls])`
//                        -> H[label][geometric word], a joint histogram of (label, GW assignment)
//
// HBM-bound integer work: 8 bytes per point, read once.  Each CTA keeps a private histogram in shared memory (32-bit
// shared atomics), streams its grid-stride share of the points with 16-byte loads and flushes the non-zero bins to the
// 64-bit global result with one atomic per bin.  Points whose labels fall outside [0, NA) x [0, NB) are skipped (255 =
// "ignore" in the loaders).  Exact: integer counts, any summation order.
#include "common.cuh"

namespace gfs {

constexpr int JH_THREADS = 256;
constexpr int JH_MAXBINS = 12288;   // 48 KiB of 32-bit bins

__global__ void __launch_bounds__(JH_THREADS)
joint_histogram_kernel(const int32_t* __restrict__ a, const int32_t* __restrict__ b, int64_t n, int NA, int NB,
                       unsigned long long* __restrict__ out) {
    extern __shared__ uint32_t bins[];
    const int nb = NA * NB;
    for (int i = threadIdx.x; i < nb; i += JH_THREADS) bins[i] = 0u;
    __syncthreads();
    auto add = [&](int32_t x, int32_t y) {
        if ((unsigned)x < (unsigned)NA && (unsigned)y < (unsigned)NB) atomicAdd(&bins[x * NB + y], 1u);
    };
    const int64_t n4 = n >> 2;
    const int4* a4 = reinterpret_cast<const int4*>(a);
    const int4* b4 = reinterpret_cast<const int4*>(b);
    for (int64_t i = (int64_t)blockIdx.x * JH_THREADS + threadIdx.x; i < n4; i += (int64_t)gridDim.x * JH_THREADS) {
        const int4 u = __ldg(a4 + i), v = __ldg(b4 + i);
        add(u.x, v.x);
        add(u.y, v.y);
        add(u.z, v.z);
        add(u.w, v.w);
    }
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += JH_THREADS) add(a[i], b[i]);
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += JH_THREADS) {
        const uint32_t c = bins[i];
        if (c) atomicAdd(out + i, (unsigned long long)c);
    }
}

}  // namespace gfs

extern "C" int gfs_joint_histogram_i32(const int32_t* a, const int32_t* b, int64_t n, int NA, int NB, unsigned long long* counts,
                                       void* stream) {
    using namespace gfs;
    GFS_REQUIRE(a && b && counts, GFS_ERR_BAD_ARG, "gfs_joint_histogram_i32: null pointer");
    GFS_REQUIRE(n >= 0 && NA > 0 && NB > 0, GFS_ERR_BAD_ARG, "gfs_joint_histogram_i32: bad sizes (n=%lld NA=%d NB=%d)", (long long)n, NA, NB);
    GFS_REQUIRE((int64_t)NA * NB <= JH_MAXBINS, GFS_ERR_UNSUPPORTED, "gfs_joint_histogram_i32: %d x %d bins exceed %d", NA, NB, JH_MAXBINS);
    GFS_REQUIRE((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0, GFS_ERR_BAD_ARG,
                "gfs_joint_histogram_i32: inputs must be 16-byte aligned");
    if (n == 0) return GFS_OK;
    const int sms = sm_count();
    GFS_REQUIRE(sms > 0, GFS_ERR_CUDA, "gfs_joint_histogram_i32: cannot query the device");
    // every point contributes < 2^32 per CTA bin only if a CTA sees < 2^32 points: cap the share per CTA
    int64_t grid = (n / 4 + JH_THREADS - 1) / JH_THREADS;
    const int64_t cap = (int64_t)sms * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    GFS_REQUIRE(n / grid < ((int64_t)1 << 32), GFS_ERR_UNSUPPORTED, "gfs_joint_histogram_i32: n=%lld too large for 32-bit CTA bins", (long long)n);
    const size_t smem = (size_t)NA * NB * sizeof(uint32_t);
    joint_histogram_kernel<<<(unsigned)grid, JH_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(a, b, n, NA, NB, counts);
    GFS_LAUNCH_OK("joint_histogram_kernel");
    return GFS_OK;
}
