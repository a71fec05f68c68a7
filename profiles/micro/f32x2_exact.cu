// Are add / sub / mul .rn.f32x2 bit-identical to the scalar instructions?  (random bit patterns incl. subnormals, NaN, inf)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
__global__ void k(unsigned long long* bad, uint32_t* example) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  for (int i = 0; i < 4096; ++i) {
    uint32_t b[4];
    for (int q = 0; q < 4; ++q) {
      b[q] = rng(s);
      const uint32_t mode = rng(s) & 15u;
      if (mode == 0) b[q] &= 0x807fffffu;                       // subnormal
      else if (mode == 1) b[q] = (b[q] & 0x80000000u);           // signed zero
      else if (mode < 8) b[q] = (b[q] & 0x807fffffu) | ((100u + (rng(s) % 56u)) << 23);  // moderate exponents
    }
    const float a0 = __uint_as_float(b[0]), a1 = __uint_as_float(b[1]), c0 = __uint_as_float(b[2]), c1 = __uint_as_float(b[3]);
    float r0, r1, s0, s1;
    for (int op = 0; op < 3; ++op) {
      if (op == 0) {
        asm volatile("{.reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; add.rn.f32x2 c, a, b; mov.b64 {%0, %1}, c;}" : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(c0), "f"(c1));
        asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(s0) : "f"(a0), "f"(c0));
        asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(s1) : "f"(a1), "f"(c1));
      } else if (op == 1) {
        asm volatile("{.reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; sub.rn.f32x2 c, a, b; mov.b64 {%0, %1}, c;}" : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(c0), "f"(c1));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(s0) : "f"(a0), "f"(c0));
        asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(s1) : "f"(a1), "f"(c1));
      } else {
        asm volatile("{.reg .b64 a, b, c; mov.b64 a, {%2, %3}; mov.b64 b, {%4, %5}; mul.rn.f32x2 c, a, b; mov.b64 {%0, %1}, c;}" : "=f"(r0), "=f"(r1) : "f"(a0), "f"(a1), "f"(c0), "f"(c1));
        asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(s0) : "f"(a0), "f"(c0));
        asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(s1) : "f"(a1), "f"(c1));
      }
      const bool nan0 = s0 != s0 && r0 != r0, nan1 = s1 != s1 && r1 != r1;
      const bool bad0 = !nan0 && __float_as_uint(r0) != __float_as_uint(s0), bad1 = !nan1 && __float_as_uint(r1) != __float_as_uint(s1);
      if (bad0 || bad1) {
        if (atomicAdd(bad + op, 1ull) == 0) {
          example[op * 6 + 0] = bad0 ? b[0] : b[1]; example[op * 6 + 1] = bad0 ? b[2] : b[3];
          example[op * 6 + 2] = __float_as_uint(bad0 ? r0 : r1); example[op * 6 + 3] = __float_as_uint(bad0 ? s0 : s1);
        }
      }
    }
  }
}
int main() {
  unsigned long long* bad; uint32_t* ex;
  cudaMallocManaged(&bad, 3 * 8); cudaMallocManaged(&ex, 18 * 4);
  for (int i = 0; i < 3; ++i) bad[i] = 0;
  k<<<1024, 256>>>(bad, ex);
  cudaDeviceSynchronize();
  const char* n[] = {"add", "sub", "mul"};
  for (int i = 0; i < 3; ++i) printf("%s.rn.f32x2 vs scalar: %llu mismatches of %llu; first: a=%08x b=%08x packed=%08x scalar=%08x\n", n[i], bad[i], 1024ull * 256 * 4096 * 2, ex[i*6], ex[i*6+1], ex[i*6+2], ex[i*6+3]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
