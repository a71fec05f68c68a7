// s2d_math.cuh - device math of the lockstep soccer simulator (sm_100a).
//
// fp32 SPEC.  Every value the simulator produces is defined by a fixed sequence of IEEE-754 binary32
// operations: +, -, *, /, sqrt (all round-to-nearest: nvcc's default -prec-div/-prec-sqrt), rint, and
// fused multiply-adds ONLY where this file writes __fmaf_rn.  The translation unit is compiled with
// --fmad=false, so ptxas never contracts a*b+c on its own.  That makes results independent of compiler
// scheduling and reproducible on any IEEE machine (the test suite re-states the same sequence in C).
//
// Angles are DEGREES in [-180, 180] as in the proto fields (idl/service.proto:181-223) and in pyrusgeom's
// AngleDeg (call sites sample_environments/reach_ball_env.py:89-96, :119-124).
//
// S2D_HOST_EMU: tests/emu compiles the __device__ functions of csrc/ for the host (with a shim header for
// the few CUDA built-ins) to check them bit for bit on a machine without a GPU.  The product library never
// defines it.
#pragma once
#ifndef S2D_HOST_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace s2d {

// min / max: FMNMX.  (NaN never occurs on valid states; the sign of a zero result is not significant.)
__device__ __forceinline__ float fmin_(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ float clampf(float lo, float x, float hi) { return fmaxf(lo, fminf(x, hi)); }

__device__ __forceinline__ float hypot2(float x, float y) { return sqrtf(x * x + y * y); }
// The same value for an operand that is often the zero vector (the velocity of a player at rest): sqrt(0) is answered
// by the IEEE square root's out-of-line slow path, and a call makes the warp wait for every load in flight first.
__device__ __forceinline__ float hypot2_or_zero(float x, float y) {
  const bool zero = x == 0.0f && y == 0.0f;
  const float h = hypot2(zero ? 1.0f : x, y);
  return zero ? 0.0f : h;
}

// cold paths kept out of line: they are (almost) never taken, and inlining them costs registers and I-cache
__device__ __noinline__ float cold_fmod360(float d) { return fmodf(d, 360.0f); }
__device__ __noinline__ float cold_div(float a, float b) { return a / b; }

// AngleDeg normalisation: fmod only beyond +-360, then one wrap.  Both ends of [-180, 180] are kept.
__device__ __forceinline__ float norm_deg(float d) {
  if (d < -360.0f || 360.0f < d) d = cold_fmod360(d);
  if (d < -180.0f) d += 360.0f;
  if (d > 180.0f) d -= 360.0f;
  return d;
}
// the same for |d| <= 360 (sum or difference of two normalised angles): the fmod is never reached
__device__ __forceinline__ float norm_deg_360(float d) {
  d = d < -180.0f ? d + 360.0f : d;
  d = d > 180.0f ? d - 360.0f : d;
  return d;
}

// sin and cos of x degrees.  Quadrant reduction is exact in degrees (q*90 is exact, the fma has a single
// rounding of an exactly representable result); the kernels are the classic single-precision minimax
// polynomials for sin and cos on [-pi/4, pi/4].
__device__ __forceinline__ void sincos_deg(float x, float& s, float& c) {
  const float q = rintf(x * 0.011111111f);
  const float r = __fmaf_rn(q, -90.0f, x);
  const float t = r * 0.017453292f;
  const float z = t * t;
  float u = __fmaf_rn(z, -1.9515295891e-4f, 8.3321608736e-3f);
  u = __fmaf_rn(z, u, -1.6666654611e-1f);
  const float sp = __fmaf_rn(t * z, u, t);
  float v = __fmaf_rn(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  v = __fmaf_rn(z, v, 4.166664568298827e-2f);
  const float cp = __fmaf_rn(z * z, v, __fmaf_rn(z, -0.5f, 1.0f));
  const int n = static_cast<int>(q) & 3;
  float ss = (n & 1) ? cp : sp;
  float cc = (n & 1) ? sp : cp;
  if (n == 1 || n == 2) cc = -cc;
  if (n >= 2) ss = -ss;
  s = ss;
  c = cc;
}

// Memo of sincos_deg over the whole degrees in [-360, 360] ({sin, cos} per entry, 5.8 KB in global memory, read
// through L1 by the ReachBall Discrete(n) step kernel).  A player placed at a whole-degree body angle that only
// dashes in directions snapped to dash_angle_step = 1 asks for nothing else, and so do the ball-velocity draws of a
// placement; a table read is a quarter of the polynomial's instructions.  The
// entries are written by sincos_deg itself (fill_sincos_memo, s2d_api.cu), so a hit returns the very bits the
// polynomial would; any other argument takes the polynomial, out of line.
constexpr int kSinCosMemoHalf = 360, kSinCosMemoSize = 2 * kSinCosMemoHalf + 1;
__device__ __noinline__ float2 cold_sincos_deg(float x) {
  float2 r;
  sincos_deg(x, r.x, r.y);
  return r;
}
__device__ __forceinline__ void sincos_deg_memo(float x, const float2* __restrict__ memo, float& s, float& c) {
  const float r = rintf(x);
  if (r == x && fabsf(x) <= static_cast<float>(kSinCosMemoHalf)) {
    const float2 t = __ldg(memo + (static_cast<int>(r) + kSinCosMemoHalf));
    s = t.x;
    c = t.y;
  } else {
    const float2 t = cold_sincos_deg(x);
    s = t.x;
    c = t.y;
  }
}

// Vector2D::th() in degrees: octant reduction, ONE division, odd minimax polynomial; 0 for the zero vector.
__device__ __forceinline__ float atan2_deg(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const bool swap = ay > ax;
  const float mx = swap ? ay : ax, mn = swap ? ax : ay;
  const bool upper = mn > 0.41421356f * mx;
  const float num = upper ? mn - mx : mn;
  float den = upper ? mn + mx : mx;
  const float off = upper ? 45.0f : 0.0f;
  const bool zero = mx == 0.0f;  // both components are zero
  den = zero ? 1.0f : den;       // 0 / 1 = 0 -> polynomial gives 0 -> result 0 (or +-180 / +-0 folded below)
  const float a = num / den;
  const float z = a * a;
  float p = __fmaf_rn(z, 8.05374449538e-2f, -1.38776856032e-1f);
  p = __fmaf_rn(z, p, 1.99777106478e-1f);
  p = __fmaf_rn(z, p, -3.33329491539e-1f);
  p = __fmaf_rn(p * z, a, a);
  float r = __fmaf_rn(p, 57.29577951f, off);
  r = swap ? 90.0f - r : r;
  r = x < 0.0f ? 180.0f - r : r;
  r = y < 0.0f ? -r : r;
  return zero ? 0.0f : r;
}

// e^x, |x| <= 88: Cody-Waite reduction by ln2 and a degree-6 polynomial, exponent patched in.
__device__ __forceinline__ float exp_poly(float x) {
  const float k = rintf(x * 1.44269504f);
  float r = __fmaf_rn(k, -0.693359375f, x);
  r = __fmaf_rn(k, 2.12194440e-4f, r);
  float p = __fmaf_rn(r, 1.9875691500e-4f, 1.3981999507e-3f);
  p = __fmaf_rn(r, p, 8.3334519073e-3f);
  p = __fmaf_rn(r, p, 4.1665795894e-2f);
  p = __fmaf_rn(r, p, 1.6666665459e-1f);
  p = __fmaf_rn(r, p, 5.0000001201e-1f);
  p = __fmaf_rn(p, r * r, r) + 1.0f;
  // (the exponent goes in through unsigned arithmetic: shifting a negative int is undefined before C++20)
  return __uint_as_float(__float_as_uint(p) + (static_cast<uint32_t>(static_cast<int>(k)) << 23));
}

// softmax([a, b])[0] = e^a / (e^a + e^b), written with one exponential
__device__ __forceinline__ float softmax_first(float a, float b) { return 1.0f / (1.0f + exp_poly(b - a)); }

// ------------------------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
// 1, 2, 3", SC'11).  key = the 64-bit seed; counter = (global env id lo, hi, index, purpose<<24 | sub).
// `index` is the episode number for reset draws and the server cycle for per-step draws, so a stream never
// depends on how envs are sharded over GPUs or how many substeps a launch fuses.
// ------------------------------------------------------------------------------------------------------
enum RngPurpose : uint32_t { RNG_RESET = 0, RNG_BALLVEL = 1, RNG_ACTION = 2, RNG_NOISE = 3, RNG_TACKLE = 5 };  // (4: player-type draws)

__device__ __forceinline__ uint4 philox4x32_10(uint64_t seed, uint64_t env, uint32_t index, uint32_t purpose,
                                               uint32_t sub) {
  uint32_t c0 = static_cast<uint32_t>(env), c1 = static_cast<uint32_t>(env >> 32), c2 = index,
           c3 = (purpose << 24) | sub;
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0;
    c1 = l1;
    c2 = h0 ^ c3 ^ k1;
    c3 = l0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// uniform integer in [lo, hi] (multiply-shift) - the role of random.randint at reach_ball_env.py:173-178
__device__ __forceinline__ int u32_to_int(uint32_t u, int lo, int hi) {
  return lo + static_cast<int>(__umulhi(u, static_cast<uint32_t>(hi - lo + 1)));
}
// uniform in [0, 1), 24 bits - the role of random.random at reach_ball_env.py:205
__device__ __forceinline__ float u32_to_unit(uint32_t u) { return static_cast<float>(u >> 8) * (1.0f / 16777216.0f); }

#ifndef S2D_HOST_EMU
// 128-bit streaming accesses: state planes are touched exactly once per launch, so do not allocate in L1.
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
// 256-bit accesses (sm_100): one full 32-byte sector per lane
__device__ __forceinline__ void ld_nc_256(const float4* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st_stream_256(float4* p, const float4& a, const float4& b) {
  asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z),
               "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}
// asks the L2 for the line of p: one instruction per lane, no registers held while the data travels
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
#else
__device__ __forceinline__ float ld_stream(const float* p) { return *p; }
__device__ __forceinline__ void st_stream(float* p, float v) { *p = v; }
__device__ __forceinline__ float4 ld_stream(const float4* p) { return *p; }
__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return *p; }
__device__ __forceinline__ void st_stream(float4* p, const float4& v) { *p = v; }
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) { *p = v; }
#endif

}  // namespace s2d
