// ubench2.cu -- clean IMAD.WIDE.U32 peak probe: loop body = 64 IMAD.WIDE.U32 and nothing else.
// acc[k] = lo32(acc[(k+1)%8]) * b + acc[k]   (data dependent: cannot be hoisted or strength-reduced)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
template <int WARPS_PER_BLOCK>
__global__ void probe(int iters, u32* sink, u32 b) {
  const u32 seed = blockIdx.x * blockDim.x + threadIdx.x + 1u;
  u64 a0 = seed, a1 = seed * 3u, a2 = seed * 5u, a3 = seed * 7u, a4 = seed * 11u, a5 = seed * 13u, a6 = seed * 17u, a7 = seed * 19u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a0) : "r"((u32)a1), "r"(b));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a1) : "r"((u32)a2), "r"(b));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a2) : "r"((u32)a3), "r"(b));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a3) : "r"((u32)a4), "r"(b));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a4) : "r"((u32)a5), "r"(b));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a5) : "r"((u32)a6), "r"(b));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a6) : "r"((u32)a7), "r"(b));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a7) : "r"((u32)a0), "r"(b));
    }
  }
  u64 s = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (s == 0x1234567ull) sink[0] = (u32)s;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double mhz = p.clockRate / 1000.0;
  u32* sink; cudaMalloc(&sink, 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int wps = 1; wps <= 16; wps *= 2) {          // warps per scheduler
    int threads = 128, blocks = sms * wps;            // blocks of 4 warps: wps blocks per SM -> wps warps per scheduler
    int iters = 4000;
    probe<4><<<blocks, threads>>>(100, sink, 12345u); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(a); probe<4><<<blocks, threads>>>(iters, sink, 12345u + r); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    double ops = 64.0 * iters * (double)blocks * threads;
    printf("warps/scheduler %2d : %.3e IMAD.WIDE/s = %.2f /clk/SM @%.0f MHz\n", wps, ops / (best * 1e-3), ops / (best * 1e-3) / sms / (mhz * 1e6), mhz);
  }
  return 0;
}
