// micro-benchmark: tcgen05.ld throughput per SM as a function of the number of warps and the load shape
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/ldtm_bw scripts/micro/ldtm_bw.cu ; run on the B200
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ uint32_t ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld<32>(uint32_t taddr) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= r[i];
    return x;
}
template <>
__device__ __forceinline__ uint32_t ld<8>(uint32_t taddr) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\ntcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
    return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}

template <int X>
__global__ void k(int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) acc ^= ld<X>(t + ((i * X + (warp >> 2) * 64) & 511 & ~(X - 1)));
    __syncthreads();
    const long long c1 = clock64();
    if (threadIdx.x == 0) *cycles = c1 - c0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tbase) : "memory");
}

int main() {
    long long* cyc;
    uint32_t* sink;
    cudaMalloc(&cyc, 8);
    cudaMalloc(&sink, 4 * 1024 * 148);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32}) {
        for (int x : {8, 32}) {
            long long h = 0;
            if (x == 32) k<32><<<1, warps * 32>>>(iters, cyc, sink); else k<8><<<1, warps * 32>>>(iters, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double bytes = (double)iters * warps * 32 * x * 4;
            printf("warps %2d shape 32x32b.x%-2d : %lld cycles, %.1f B/cycle/SM, %.1f cycles per load  (%s)\n", warps, x, h, bytes / h, (double)h / iters,
                   cudaGetErrorString(e));
        }
    }
    return 0;
}
