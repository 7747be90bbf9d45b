// Microbenchmark: TMEM -> register bandwidth of tcgen05.ld on sm_100a (decides the attention softmax design: how many
// times a score tile may be read).  One CTA per SM, 512 TMEM columns, W warps each issuing `iters` tcgen05.ld.32x32b.xN
// back to back (one tcgen05.wait::ld per `batch` loads); reports bytes / clock / SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/tmem_bw.bin tools/microbench/tmem_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t& sink);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t& sink) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) sink ^= r[i];
}
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t& sink) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) sink ^= r[i];
}

// NOWAIT variant: 4 x32 loads in flight, then one wait (what the attention kernel does for a 128-column row)
__device__ __forceinline__ void ld128(uint32_t taddr, uint32_t& sink) {
  uint32_t r[128];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t* q = r + c * 32;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
          "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
          "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
        : "r"(taddr + c * 32)
        : "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 128; ++i) sink ^= r[i];
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_bw_kernel(int iters, long long* clocks, uint32_t* sinkp) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tmem_base + ((static_cast<uint32_t>(warp & 3) * 32) << 16);
  uint32_t sink = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t col = static_cast<uint32_t>((i * 128) & 511) ;
    if (MODE == 0) { ld<32>(t + (col & 480), sink); }
    else if (MODE == 1) { ld<16>(t + (col & 496), sink); }
    else { ld128(t + (col & 384), sink); }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  if (sink == 0x12345678u) sinkp[0] = sink;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* clocks;
  uint32_t* sink;
  cudaMalloc(&clocks, sms * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int iters = 2048;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 2, 4, 8, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) tmem_bw_kernel<0><<<sms, warps * 32>>>(iters, clocks, sink);
        else if (mode == 1) tmem_bw_kernel<1><<<sms, warps * 32>>>(iters, clocks, sink);
        else tmem_bw_kernel<2><<<sms, warps * 32>>>(iters, clocks, sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      }
      long long c = 0;
      cudaMemcpy(&c, clocks, sizeof(c), cudaMemcpyDeviceToHost);
      const double bytes_per_ld = (mode == 0 ? 32 : mode == 1 ? 16 : 128) * 32 * 4.0;
      printf("%s  warps/CTA %2d: %8lld clk for %d loads per warp -> %.1f B/clk per warp, %.1f B/clk per SM\n",
             mode == 0 ? "32x32b.x32 + wait   " : mode == 1 ? "32x32b.x16 + wait   " : "4 x (x32) then wait ", warps, c, iters,
             bytes_per_ld * iters / c, bytes_per_ld * iters * warps / c);
    }
  return 0;
}
