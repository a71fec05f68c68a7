// Why the kernels do not use sm_100's packed fp32 instructions (add / mul .rn.f32x2 = FADD2 / FMUL2), although they
// halve the issue slots of (x, y) arithmetic (f32x2.cu: same arithmetic throughput, one issue slot for two operations;
// move_object written with them made the K = 16 step kernel 1.6 % faster):
// ptxas CONTRACTS mul.rn.f32x2 followed by add.rn.f32x2 into FFMA2 - even with an explicit .rn on both and --fmad=false
// (nvcc passes "--fmad false" to ptxas) - so results differ in the last bit from the scalar sequence the specification
// and the CPU oracle use (the noise tests caught it).  Each packed instruction alone is bit-identical to its scalar
// form (f32x2_exact.cu: 0 mismatches in 6.4e9 operations incl. subnormals).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o fuse f32x2_fusion.cu && cuobjdump -sass fuse | grep FFMA2 && ./fuse
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tadd.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 a, b, c;\n\tmov.b64 a, {%2, %3};\n\tmov.b64 b, {%4, %5};\n\tmul.rn.f32x2 c, a, b;\n\tmov.b64 {%0, %1}, c;}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__global__ void k(unsigned long long* bad, float* ex) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 77u;
  auto rng = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
  for (int i = 0; i < 1024; ++i) {
    const float vx = (int(rng() % 2001) - 1000) * 1e-3f, vy = (int(rng() % 2001) - 1000) * 1e-3f;
    const float ux = (int(rng() % 2001) - 1000) * 1e-3f, uy = (int(rng() % 2001) - 1000) * 1e-3f, m = (rng() % 1000) * 1e-4f;
    float sx = vx, sy = vy;
    sx += m * ux;  // (--fmad=false: FMUL, FADD)
    sy += m * uy;
    const float2 p = add2(make_float2(vx, vy), mul2(make_float2(ux, uy), make_float2(m, m)));
    if (__float_as_uint(p.x) != __float_as_uint(sx) || __float_as_uint(p.y) != __float_as_uint(sy))
      if (atomicAdd(bad, 1ull) == 0) { ex[0] = vx; ex[1] = m; ex[2] = ux; ex[3] = p.x; ex[4] = sx; }
  }
}
int main() {
  unsigned long long* bad; float* ex;
  cudaMallocManaged(&bad, 8); cudaMallocManaged(&ex, 64);
  *bad = 0;
  k<<<256, 256>>>(bad, ex);
  cudaDeviceSynchronize();
  printf("v + m * u, packed against scalar: %llu mismatches of %d; first: v=%.9g m=%.9g u=%.9g packed=%.9g scalar=%.9g  (%s)\n",
         *bad, 256 * 256 * 1024, ex[0], ex[1], ex[2], ex[3], ex[4], cudaGetErrorString(cudaGetLastError()));
}
