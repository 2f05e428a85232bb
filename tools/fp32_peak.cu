// FP32-pipe microbenchmark for the PP roofline denominator ("measure with an FMA microbenchmark") and for the
// choice between scalar and packed (f32x2) arithmetic in the FFT butterflies.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp32_peak tools/fp32_peak.cu && ./fp32_peak
// Prints one JSON object: achieved TFLOP/s (2 flop per FMA, 1 per ADD) of FFMA, FFMA2, FADD, FADD2 with all SMs busy.
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE> __global__ void __launch_bounds__(256) k(float* out, float s) {
  float2 a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f - i);
  const float2 m = make_float2(s, s * 0.5f), c = make_float2(1e-3f, -1e-3f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
      if (MODE == 1) a[i] = __ffma2_rn(a[i], m, c);
      if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
      if (MODE == 3) a[i] = __fadd2_rn(a[i], c);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) r += a[i].x + a[i].y;
  if (r == 12345.678f) out[0] = r;
}

template <int MODE> double run(float* d, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(d, 0.999f);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, 0.999f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double ops = (double)blocks * 256 * ITERS * ILP * 2;   // scalar-equivalent operations
  const double flop = ops * ((MODE < 2) ? 2.0 : 1.0);
  return flop / (best * 1e-3) / 1e12;
}

int main() {
  float* d; cudaMalloc(&d, 4);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int blocks = p.multiProcessorCount * 8;
  const double a = run<0>(d, blocks), b = run<1>(d, blocks), c = run<2>(d, blocks), e = run<3>(d, blocks);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"fadd_tflops\": %.2f, \"fadd2_tflops\": %.2f}\n", p.name,
         p.multiProcessorCount, a, b, c, e);
  return 0;
}
