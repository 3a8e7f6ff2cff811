// Micro-benchmarks that bound the fused ISP sweep on B200 (run through gpurun; results -> profiles/).
//   1. achievable HBM bandwidth for the sweep's traffic mix (1.5 B read : 6 B written per pixel) with the
//      sweep's thread mapping (thread = 8 pixels of a row, warp = 256-pixel strip walking down the rows),
//      for different store shapes: per-thread 3 x 16 B (48-byte lane stride), warp-transposed through
//      shared memory (each STG.128 covers 512 contiguous bytes), plain copy 1:1, pure read, pure write.
//   2. issue rate of the packed FP32 instructions (FFMA2 / FADD2) against scalar FFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int H = 3648, W = 5472, NF = 6;
constexpr int PITCH_W = W * 3 / 8;        // words per packed row
constexpr int NTCOLS = W / 8;             // 684 thread columns
constexpr int WPR = (NTCOLS + 31) / 32;   // 22 warps per row

// MODE 0: direct 3 x uint4 per thread; 1: smem transpose, 3 x uint4 per lane contiguous per warp; 2: read only; 3: write only (direct)
// 4: write only (contiguous); 5: direct with st.global.cs; 6: transposed + st.cs
template <int MODE, int BLOCKS_PER_SM>
__global__ void __launch_bounds__(256, BLOCKS_PER_SM) mix_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                int rows_per_task, int nchunks, long long total_tasks, uint32_t* sink) {
  __shared__ uint4 tile[8][3 * 32 + 3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  long long task = (long long)blockIdx.x * 8 + wid;
  if (task >= total_tasks) return;
  const int strip = (int)(task % WPR);
  const long long t2 = task / WPR;
  const int chunk = (int)(t2 % nchunks), frame = (int)(t2 / nchunks);
  const int tc = strip * 32 + lane;
  const bool act = tc < NTCOLS;
  const int tcol = act ? tc : NTCOLS - 1;
  const int r0 = chunk * rows_per_task, rend = min(r0 + rows_per_task, H);
  const uint32_t* src = in + (size_t)frame * H * PITCH_W + 3 * tcol;
  uint32_t* dst = out + (size_t)frame * H * W * 3 / 2 + 12 * tcol;         // u16 RGB: 12 words per 8 px
  uint32_t* wdst = out + (size_t)frame * H * W * 3 / 2 + 12 * (strip * 32);  // warp base
  const int nvalid = min(32, NTCOLS - strip * 32);                        // lanes with data in this strip
  uint32_t acc = 0;
  uint32_t a0, a1, a2, b0, b1, b2;
  if (MODE != 3 && MODE != 4) {
    a0 = __ldg(src + (size_t)r0 * PITCH_W); a1 = __ldg(src + (size_t)r0 * PITCH_W + 1); a2 = __ldg(src + (size_t)r0 * PITCH_W + 2);
    b0 = __ldg(src + (size_t)(r0 + 1) * PITCH_W); b1 = __ldg(src + (size_t)(r0 + 1) * PITCH_W + 1); b2 = __ldg(src + (size_t)(r0 + 1) * PITCH_W + 2);
  } else { a0 = a1 = a2 = b0 = b1 = b2 = tcol; }
#pragma unroll 1
  for (int row = r0; row < rend; row += 2) {
    uint32_t c0 = a0, c1 = a1, c2 = a2, d0 = b0, d1 = b1, d2 = b2;
    if (MODE != 3 && MODE != 4) {
      const int rn = min(row + 2, H - 2);
      a0 = __ldg(src + (size_t)rn * PITCH_W); a1 = __ldg(src + (size_t)rn * PITCH_W + 1); a2 = __ldg(src + (size_t)rn * PITCH_W + 2);
      b0 = __ldg(src + (size_t)(rn + 1) * PITCH_W); b1 = __ldg(src + (size_t)(rn + 1) * PITCH_W + 1); b2 = __ldg(src + (size_t)(rn + 1) * PITCH_W + 2);
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const uint32_t x0 = rr ? d0 : c0, x1 = rr ? d1 : c1, x2 = rr ? d2 : c2;
      if (MODE == 2) { acc += x0 ^ x1 ^ x2; continue; }
      uint4 q[3];
      q[0] = make_uint4(x0, x1 + 1, x2 + 2, x0 + 3);
      q[1] = make_uint4(x1 + 4, x2 + 5, x0 + 6, x1 + 7);
      q[2] = make_uint4(x2 + 8, x0 + 9, x1 + 10, x2 + 11);
      const size_t roff = (size_t)(row + rr) * (W * 3 / 2);
      if (MODE == 0 || MODE == 3 || MODE == 5) {
        if (act) {
          uint4* d = reinterpret_cast<uint4*>(dst + roff);
          if (MODE == 5) { __stcs(d, q[0]); __stcs(d + 1, q[1]); __stcs(d + 2, q[2]); }
          else { d[0] = q[0]; d[1] = q[1]; d[2] = q[2]; }
        }
      } else {
        __syncwarp();
        tile[wid][3 * lane] = q[0]; tile[wid][3 * lane + 1] = q[1]; tile[wid][3 * lane + 2] = q[2];
        __syncwarp();
        uint4* d = reinterpret_cast<uint4*>(wdst + roff);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int idx = i * 32 + lane;
          if (idx < 3 * nvalid) {
            if (MODE == 6) __stcs(d + idx, tile[wid][idx]); else d[idx] = tile[wid][idx];
          }
        }
      }
    }
  }
  if (MODE == 2 && acc == 0x12345u) *sink = acc;
}

__global__ void copy_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    out[i] = a; out[i + stride] = b; out[i + 2 * stride] = c; out[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) out[i] = in[i];
}

template <int KIND>   // 0 FFMA, 1 FFMA2, 2 FADD2+FFMA2 mix, 3 FFMA + LOP3 interleaved, 4 FFMA2 + LOP3 interleaved
__global__ void __launch_bounds__(256) fp_kernel(float* out, float a, float b, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = a + i + threadIdx.x;
  unsigned long long* xp = reinterpret_cast<unsigned long long*>(x);
  unsigned long long ab;
  { float2 t = make_float2(a, b); ab = *reinterpret_cast<unsigned long long*>(&t); }
  uint32_t m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) m[i] = threadIdx.x * (i + 3);
  for (int it = 0; it < iters; ++it) {
    if (KIND == 0 || KIND == 3) {
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(xp[i]) : "l"(ab));
      if (KIND == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(xp[i]) : "l"(ab));
      }
    }
    if (KIND >= 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) m[i] = (m[i] & 0x7ff800u) | (m[(i + 1) & 7] >> 3);
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

template <int MODE, int BPS> void run_mix(const char* name, const uint32_t* in, uint32_t* out, uint32_t* sink, int rpt) {
  const int nchunks = (H + rpt - 1) / rpt;
  const long long tasks = (long long)NF * nchunks * WPR;
  const unsigned blocks = (unsigned)((tasks + 7) / 8);
  float ms = time_ms([&] { mix_kernel<MODE, BPS><<<blocks, 256>>>(in, out, rpt, nchunks, tasks, sink); }, 20);
  CK(cudaGetLastError());
  const double px = (double)NF * H * W;
  double bytes = (MODE == 2) ? px * 1.5 : (MODE == 3 || MODE == 4) ? px * 6 : px * 7.5;
  printf("%-44s rpt %3d blocks/SM %d : %.3f ms  %.0f GB/s\n", name, rpt, BPS, ms, bytes / ms / 1e6);
}

int main() {
  const size_t in_bytes = (size_t)NF * H * W * 3 / 2, out_bytes = (size_t)NF * H * W * 6;
  uint32_t *in, *out, *sink;
  CK(cudaMalloc(&in, in_bytes)); CK(cudaMalloc(&out, out_bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(in, 0x5a, in_bytes)); CK(cudaMemset(out, 0, out_bytes));
  {
    const size_t n = out_bytes / 2 / 16;
    uint4* a = reinterpret_cast<uint4*>(out); uint4* b = a + n;
    for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
      float ms = time_ms([&] { copy_kernel<<<blocks, 256>>>(a, b, n); }, 20);
      printf("copy 1:1 (uint4 x4 grid-stride, %d blocks): %.3f ms  %.0f GB/s\n", blocks, ms, 2.0 * n * 16 / ms / 1e6);
    }
    float ms = time_ms([&] { CK(cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice)); }, 20);
    printf("cudaMemcpy D2D: %.3f ms  %.0f GB/s\n", ms, 2.0 * n * 16 / ms / 1e6);
    ms = time_ms([&] { CK(cudaMemsetAsync(out, 1, out_bytes)); }, 20);
    printf("cudaMemset (pure write): %.3f ms  %.0f GB/s\n", ms, (double)out_bytes / ms / 1e6);
  }
  for (int rpt : {24, 48, 96}) {
    run_mix<2, 2>("read only (packed rows)", in, out, sink, rpt);
    run_mix<3, 2>("write only, per-thread 3x16B", in, out, sink, rpt);
    run_mix<4, 2>("write only, warp-contiguous via smem", in, out, sink, rpt);
    run_mix<0, 2>("mix 1.5:6, per-thread 3x16B", in, out, sink, rpt);
    run_mix<1, 2>("mix 1.5:6, warp-contiguous via smem", in, out, sink, rpt);
    run_mix<5, 2>("mix 1.5:6, per-thread 3x16B st.cs", in, out, sink, rpt);
    run_mix<6, 2>("mix 1.5:6, warp-contiguous st.cs", in, out, sink, rpt);
    run_mix<0, 4>("mix 1.5:6, per-thread 3x16B", in, out, sink, rpt);
    run_mix<1, 4>("mix 1.5:6, warp-contiguous via smem", in, out, sink, rpt);
    run_mix<0, 8>("mix 1.5:6, per-thread 3x16B", in, out, sink, rpt);
    run_mix<1, 8>("mix 1.5:6, warp-contiguous via smem", in, out, sink, rpt);
  }
  // ---------------- FP issue rates
  float* fo; CK(cudaMalloc(&fo, 148 * 8 * 256 * 4));
  const int iters = 4096;
  auto rep = [&](const char* name, float ms, double fp_inst_per_thread_iter, double lanes) {
    const double warps = 148.0 * 8 * 8;
    const double inst = warps * iters * fp_inst_per_thread_iter;
    printf("%-28s %.3f ms  %.2f warp-inst/clk/SM (at 1965 MHz)  %.1f TFLOP/s\n", name, ms, inst / (ms * 1e-3 * 1.965e9) / 148,
           warps * 32 * iters * lanes * 2 / ms / 1e9);
  };
  rep("FFMA x16", time_ms([&] { fp_kernel<0><<<148 * 8, 256>>>(fo, 1.0001f, 0.5f, iters); }, 5), 16, 16);
  rep("FFMA2 x8", time_ms([&] { fp_kernel<1><<<148 * 8, 256>>>(fo, 1.0001f, 0.5f, iters); }, 5), 8, 16);
  rep("FFMA2 x8 + FADD2 x8", time_ms([&] { fp_kernel<2><<<148 * 8, 256>>>(fo, 1.0001f, 0.5f, iters); }, 5), 16, 24);
  rep("FFMA x16 + LOP3/SHF x16", time_ms([&] { fp_kernel<3><<<148 * 8, 256>>>(fo, 1.0001f, 0.5f, iters); }, 5), 32, 16);
  rep("FFMA2 x8 + LOP3/SHF x16", time_ms([&] { fp_kernel<4><<<148 * 8, 256>>>(fo, 1.0001f, 0.5f, iters); }, 5), 24, 16);
  CK(cudaGetLastError());
  printf("done\n");
  return 0;
}
