// Probe: the memory pattern of the 11 v 11 kernel without its arithmetic, in two state layouts.
//   layout 0 (plane-major, match-minor): plane entry of player j, match i at  plane + (j * N + i) * size   (rows 4 MB apart)
//   layout 1 (tiled): the 64 matches of a block own one contiguous tile; inside it [plane][player][64]
// Every thread = one match: per player it loads PA/PB (16 B) and PC (4 B), touches them, stores them back; loads the
// player's command (16 B of a 352-byte row) and writes its 20 bytes of the 480-byte observation row.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fg_layout fg_layout.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int NP = 22, BLK = 64;
__device__ __forceinline__ float4 ldv(const float4* p) { float4 v; asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ void stv(float4* p, float4 v) { asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
template <int LAYOUT, int OBS>
__global__ void __launch_bounds__(BLK, 10) probe(char* state, const float4* cmd, float* obs, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * BLK + threadIdx.x;
  float4 *pa, *pb; float* pc; size_t row;
  if (LAYOUT == 0) {
    pa = reinterpret_cast<float4*>(state) + i; pb = reinterpret_cast<float4*>(state + n * NP * 16) + i;
    pc = reinterpret_cast<float*>(state + n * NP * 32) + i; row = n;
  } else {
    char* tile = state + (size_t)blockIdx.x * (BLK * (NP * 36 + 80));
    pa = reinterpret_cast<float4*>(tile) + threadIdx.x; pb = reinterpret_cast<float4*>(tile + BLK * NP * 16) + threadIdx.x;
    pc = reinterpret_cast<float*>(tile + BLK * NP * 32) + threadIdx.x; row = BLK;
  }
  const float4* c = cmd + i * NP;
  float acc = 0.f;
  float4 na = ldv(pa), nb = ldv(pb); float nc = *pc; float4 ncmd = __ldg(c);
#pragma unroll 1
  for (int j = 0; j < NP; ++j) {
    float4 a = na, b = nb; float cc = nc; float4 k = ncmd;
    if (j + 1 < NP) { na = ldv(pa + (j + 1) * row); nb = ldv(pb + (j + 1) * row); nc = pc[(j + 1) * row]; ncmd = __ldg(c + j + 1); }
    a.x += k.x; a.y += k.y; b.x += k.z; cc += k.w; acc += a.x + b.y;
    stv(pa + j * row, a); stv(pb + j * row, b); pc[j * row] = cc;
    if (OBS == 1) {  // scalar stores, lanes 480 B apart
      float* o = obs + i * 120 + 4 + 5 * j;
      o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x;
    }
    if (OBS == 5) {  // eight players staged (40 floats = 5 sectors, starting half a sector into the row: floats 4 + 40 p),
                     // written by the warp together as whole sectors: 160 sectors = 5 stores, consecutive lanes = consecutive sectors
      __shared__ float stage8[BLK][41];
      float* sr = stage8[threadIdx.x] + 5 * (j & 7);
      sr[0] = a.x; sr[1] = a.y; sr[2] = a.z; sr[3] = a.w; sr[4] = b.x;
      if ((j & 7) == 7) {
        __syncwarp();
        const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;
        const int64_t first = i - lane;
        const int p = j >> 3;
        for (int s5 = 0; s5 < 5; ++s5) {
          const int piece = lane + 32 * s5, mt = piece / 5, sec = piece % 5;
          const float* rr = stage8[w0 + mt] + 8 * sec;  // (the probe ignores the half-sector offset: same traffic)
          asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(reinterpret_cast<float4*>(obs + (first + mt) * 120) + 10 * p + 2 * sec),
                       "f"(rr[0]), "f"(rr[1]), "f"(rr[2]), "f"(rr[3]), "f"(rr[4]), "f"(rr[5]), "f"(rr[6]), "f"(rr[7]) : "memory");
        }
        __syncwarp();
      }
    }
    if (OBS >= 2 && OBS <= 4) {  // four players staged in shared memory, then written as ...
      __shared__ float stage[BLK][21];
      float* sr = stage[threadIdx.x] + 5 * (j & 3);
      sr[0] = a.x; sr[1] = a.y; sr[2] = a.z; sr[3] = a.w; sr[4] = b.x;
      if ((j & 3) == 3 || j == NP - 1) {
        const int q = j >> 2;
        const float* r = stage[threadIdx.x];
        if (OBS == 2) {  // ... five 16-byte pieces per thread (the kernel before)
          float4* o = reinterpret_cast<float4*>(obs + i * 120) + 1 + 5 * q;
          for (int s = 0; s < 5; ++s) stv(o + s, make_float4(r[4 * s], r[4 * s + 1], r[4 * s + 2], r[4 * s + 3]));
        } else if (OBS == 3) {  // ... 32-byte sectors per thread (2.5 per group: emulated as 3 sector stores every group but the last half)
          float4* o = reinterpret_cast<float4*>(obs + i * 120) + 5 * q + (q & 1);
          for (int s = 0; s < 2; ++s)
            asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + 2 * s), "f"(r[8 * s]), "f"(r[8 * s + 1]), "f"(r[8 * s + 2]), "f"(r[8 * s + 3]), "f"(r[8 * s + 4]), "f"(r[8 * s + 5]), "f"(r[8 * s + 6]), "f"(r[8 * s + 7]) : "memory");
          if (q & 1) asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + 4), "f"(r[16]), "f"(r[17]), "f"(r[18]), "f"(r[19]), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]) : "memory");
        } else {  // ... by the warp together: 160 float4 pieces, consecutive lanes = consecutive pieces of a row
          __syncwarp();
          const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;
          const int64_t first = i - lane;
          for (int s = 0; s < 5; ++s) {
            const int piece = lane + 32 * s, mt = piece / 5, part = piece % 5;
            const float* rr = stage[w0 + mt] + 4 * part;
            stv(reinterpret_cast<float4*>(obs + (first + mt) * 120) + 1 + 5 * q + part, make_float4(rr[0], rr[1], rr[2], rr[3]));
          }
          __syncwarp();
        }
      }
    }
  }
  if (OBS == 1) { obs[i * 120] = acc; }
}
int main() {
  const int64_t n = 1 << 18;
  const size_t sbytes = (size_t)n * (NP * 36 + 80);
  char* state; float4* cmd; float* obs; char* flush;
  cudaMalloc(&state, sbytes); cudaMalloc(&cmd, n * NP * 16); cudaMalloc(&obs, n * 480 + 4096);  // (slack: the probe writes whole groups past the last row)
  cudaMalloc(&flush, 256 << 20);
  cudaMemset(state, 0, sbytes); cudaMemset(cmd, 0, n * NP * 16);
  if (cudaError_t e = cudaGetLastError()) { printf("setup: %s\n", cudaGetErrorString(e)); return 1; }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](auto kern, const char* name, double bytes) {
    float best = 1e9, sum = 0;
    for (int r = 0; r < 8; ++r) {
      cudaError_t c0 = cudaMemset(flush, r, 256 << 20);
      cudaError_t c1 = cudaEventRecord(e0); kern<<<n / BLK, BLK>>>(state, cmd, obs, n); cudaError_t c2 = cudaGetLastError(); cudaError_t c3 = cudaEventRecord(e1); cudaError_t c4 = cudaEventSynchronize(e1);
      float ms; cudaError_t c5 = cudaEventElapsedTime(&ms, e0, e1);
      if (r == 7 && (c0 || c1 || c2 || c3 || c4 || c5)) printf("  errors: memset %d record %d launch %d record %d sync %d elapsed %d\n", c0, c1, c2, c3, c4, c5); if (r >= 2) { sum += ms; best = ms < best ? ms : best; }
    }
    printf("%-28s mean %.1f us  best %.1f us  -> %.0f GB/s\n", name, sum / 6 * 1e3, best * 1e3, bytes / (sum / 6 * 1e-3) / 1e9);
    if (cudaError_t e = cudaGetLastError()) printf("  (%s: %s)\n", name, cudaGetErrorString(e));
  };
  const double st = 2.0 * n * NP * 36, cm = (double)n * NP * 16, ob = (double)n * 480;
  run(probe<0, 0>, "plane-major, no obs", st + cm);
  run(probe<1, 0>, "tiled, no obs", st + cm);
  run(probe<0, 1>, "plane-major, obs rows", st + cm + ob);
  run(probe<0, 2>, "obs: 16 B pieces per thread", st + cm + ob);
  run(probe<0, 3>, "obs: 32 B sectors per thread", st + cm + ob);
  run(probe<0, 4>, "obs: warp-cooperative rows", st + cm + ob);
  run(probe<0, 5>, "obs: warp-cooperative sectors", st + cm + ob * 80.0 / 120.0);  // (players 0..15 only: 2 x 8)
  return 0;
}
