// Issue-rate probe: scalar FADD / FMUL against the packed add.rn.f32x2 / mul.rn.f32x2 (FADD2 / FMUL2) of sm_100.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIter = 4096, kChains = 8;

__global__ void scalar_add(float* out, float a) {
  float v[2 * kChains];
  for (int c = 0; c < 2 * kChains; ++c) v[c] = threadIdx.x * 0.001f + c;
  for (int i = 0; i < kIter; ++i) {
#pragma unroll
    for (int c = 0; c < 2 * kChains; ++c) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(a));
  }
  float s = 0;
  for (int c = 0; c < 2 * kChains; ++c) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void packed_add(float* out, float a) {
  unsigned long long v[kChains], aa;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  for (int c = 0; c < kChains; ++c) {
    float x = threadIdx.x * 0.001f + 2 * c, y = x + 1.0f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[c]) : "f"(x), "f"(y));
  }
  for (int i = 0; i < kIter; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(aa));
  }
  float s = 0;
  for (int c = 0; c < kChains; ++c) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v[c]));
    s += x + y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void scalar_mul(float* out, float a) {
  float v[2 * kChains];
  for (int c = 0; c < 2 * kChains; ++c) v[c] = 1.0f + threadIdx.x * 0.001f + c;
  for (int i = 0; i < kIter; ++i) {
#pragma unroll
    for (int c = 0; c < 2 * kChains; ++c) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(v[c]) : "f"(a));
  }
  float s = 0;
  for (int c = 0; c < 2 * kChains; ++c) s += v[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void packed_mul(float* out, float a) {
  unsigned long long v[kChains], aa;
  asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
  for (int c = 0; c < kChains; ++c) {
    float x = 1.0f + threadIdx.x * 0.001f + 2 * c, y = x + 1.0f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[c]) : "f"(x), "f"(y));
  }
  for (int i = 0; i < kIter; ++i) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[c]) : "l"(aa));
  }
  float s = 0;
  for (int c = 0; c < kChains; ++c) {
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(v[c]));
    s += x + y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int r = 0; r < 5; ++r) launch();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  const int blocks = 148 * 8, threads = 256;
  float* out;
  cudaMalloc(&out, blocks * threads * sizeof(float));
  const double flops = double(blocks) * threads * kIter * 2 * kChains;  // scalar-equivalent operations per launch
  struct { const char* name; float ms; } r[] = {
      {"add.rn.f32   (FADD)", time_ms([&] { scalar_add<<<blocks, threads>>>(out, 1e-7f); })},
      {"add.rn.f32x2 (FADD2)", time_ms([&] { packed_add<<<blocks, threads>>>(out, 1e-7f); })},
      {"mul.rn.f32   (FMUL)", time_ms([&] { scalar_mul<<<blocks, threads>>>(out, 1.0000001f); })},
      {"mul.rn.f32x2 (FMUL2)", time_ms([&] { packed_mul<<<blocks, threads>>>(out, 1.0000001f); })},
  };
  for (auto& x : r) printf("%-22s %.3f ms  %.2f T scalar-ops/s\n", x.name, x.ms, flops / x.ms * 1e-9);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
