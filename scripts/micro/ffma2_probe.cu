#include <cuda_runtime.h>
#include <stdint.h>
__global__ void k2(float2* out, float2 a, float2 b, int n) {
  unsigned long long x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b), z = x;
  for (int i = 0; i < n; ++i) {
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(z) : "l"(x), "l"(y));
    asm volatile("add.rn.f32x2 %0, %1, %0;" : "+l"(z) : "l"(x));
    asm volatile("mul.rn.f32x2 %0, %1, %0;" : "+l"(z) : "l"(y));
  }
  out[threadIdx.x] = *reinterpret_cast<float2*>(&z);
}
