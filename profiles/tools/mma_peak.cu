// peak rate of legacy mma.sync.m16n8k16 bf16 on sm_100a: NW warps per block, 1 block per SM, 8 independent accumulators per warp
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template <int ILP>
__global__ void k(float* out, int iters) {
  float c[ILP][4];
  uint32_t a[4] = {0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u}, b0 = 0x3f803f80u, b1 = 0x3f803f80u;
#pragma unroll
  for (int i = 0; i < ILP; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456f) out[0] = s;
}
int main() {
  float* d; cudaMalloc(&d, 4);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int nw : {4, 8, 16, 32}) {
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<8><<<sms, nw * 32>>>(d, 100); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<8><<<sms, nw * 32>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flop = (double)sms * nw * iters * 8 * 2.0 * 16 * 8 * 16;
    printf("warps/SM %2d: %.1f TFLOP/s (%.3f ms)\n", nw, flop / ms / 1e9, ms);
  }
  return 0;
}
