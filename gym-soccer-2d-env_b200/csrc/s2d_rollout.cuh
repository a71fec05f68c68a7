// s2d_rollout.cuh - closed-loop rollout: the policy runs INSIDE the K loop of the step kernel.
//
// The caller of the path in the reference is SB3's DQN (dqn_stable_baselines3.py:33-55: MlpPolicy = obs -> 64 -> 64 ->
// n_actions with ReLU; `model.predict(obs)` then `env.step(action)`).  With the policy in a separate (torch) launch a
// closed loop is K = 1 by construction: obs and action go through HBM every cycle and the three small GEMMs run as
// fp32 SIMT kernels.  Here a warp evaluates the Q-network for its 32 episodes with warp-level tensor-core MMAs
// (mma.sync m16n8k8, TF32 operands, fp32 accumulate) from observations that never leave the SM, takes the greedy
// (or epsilon-greedy) action and advances the episode, K times per launch:
//
//   observation (registers) -> shared memory, row = episode           [32 x 16], features 10..15 zero
//   layer 1  [16 x 16] . [16 x 64]   per half-warp tile of 16 episodes: 2 k-steps x 8 n-tiles
//   layer 2  [16 x 64] . [64 x 64]   8 x 8; the accumulator fragment of layer 1 IS the A fragment of layer 2 when the
//   layer 3  [16 x 64] . [64 x 16]   8 x 2  weight rows are paired (2t, 2t+1) instead of (t, t+4): no shuffles
//   argmax over 16 actions inside each quad of lanes, action -> shared memory -> the lane that owns the episode
//
// Weights arrive in torch nn.Linear layout (device pointers) and are re-laid in shared memory once per block, already
// rounded to TF32, padded so that the 64-bit fragment loads are bank-conflict free.  TF32 carries 10 mantissa bits:
// Q-values agree with an fp32 evaluation to ~1e-3 relative, so the greedy action can differ on near-ties (tests).
#pragma once
#include "s2d_scenarios.cuh"

#ifndef S2D_ROLLOUT_DRAW_MEMO
#define S2D_ROLLOUT_DRAW_MEMO true
#endif
namespace s2d {

constexpr int kMlpHidden = 64;
constexpr int kMlpActions = 24;  // room for 3 column tiles of 8 actions (ReachBall uses 2: Discrete(n <= 16); Shoot 3: n <= 24)
constexpr int kMlpPad = 4;       // row padding (in float2 entries) that spreads the four t-rows over the banks

struct MlpWeights {  // device pointers, torch layout: weight [out][in] row-major, bias [out]
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  int obs_dim, n_actions;
};

// optional per-cycle output of a rollout (time-major, so that a warp's writes are contiguous): what a replay buffer needs
struct TrajOut {
  float* obs;        // [K + 1][N][10]: obs[k] = what the policy saw in cycle k, obs[K] = the final observation
  uint8_t* actions;  // [K][N] (Discrete)
  float* actions_f;  // [K][N][action_dim] (Box: the actor's rollout)
  float* reward;     // [K][N]
  uint8_t* done;     // [K][N]: the episode ended in cycle k (obs[k + 1] then opens the next episode)
};

struct MlpShared {
  float2 w1[2][4][kMlpHidden + kMlpPad];   // [k-step][t][n] = {W1[n][8k + t], W1[n][8k + t + 4]}
  float2 w2[8][4][kMlpHidden + kMlpPad];   // [k-step][t][n] = {W2[n][8k + 2t], W2[n][8k + 2t + 1]}
  float2 w3[8][4][kMlpActions + kMlpPad];  // [k-step][t][n] = {W3[n][8k + 2t], W3[n][8k + 2t + 1]}
  float b1[kMlpHidden], b2[kMlpHidden], b3[kMlpActions];
  float obs[kBlock / 32][32][20];          // row stride 20: the A-fragment loads hit 32 different banks
  uint8_t act[kBlock / 32][32];
  float actf[kBlock / 32][32][4];          // the actor's outputs (Box actions), row = episode
};

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// D = A(16x8, row) * B(8x8, col) + D.  Fragments (g = lane / 4, t = lane % 4):
//   a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4);  b0 (k = t, n = g)  b1 (k = t+4, n = g)
//   d0 (g, 2t)  d1 (g, 2t+1)  d2 (g+8, 2t)  d3 (g+8, 2t+1)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)));
}

__device__ __forceinline__ void mlp_load_weights(MlpShared& s, const MlpWeights& w, float absent_bias = -3.0e38f) {
  for (int idx = threadIdx.x; idx < 2 * 4 * kMlpHidden; idx += blockDim.x) {
    const int n = idx % kMlpHidden, t = (idx / kMlpHidden) % 4, k = idx / (4 * kMlpHidden);
    const int f0 = 8 * k + t, f1 = f0 + 4;
    s.w1[k][t][n] = make_float2(f0 < w.obs_dim ? to_tf32(__ldg(w.w1 + n * w.obs_dim + f0)) : 0.0f,
                                f1 < w.obs_dim ? to_tf32(__ldg(w.w1 + n * w.obs_dim + f1)) : 0.0f);
  }
  for (int idx = threadIdx.x; idx < 8 * 4 * kMlpHidden; idx += blockDim.x) {
    const int n = idx % kMlpHidden, t = (idx / kMlpHidden) % 4, k = idx / (4 * kMlpHidden);
    const float* row = w.w2 + n * kMlpHidden + 8 * k + 2 * t;
    s.w2[k][t][n] = make_float2(to_tf32(__ldg(row)), to_tf32(__ldg(row + 1)));
  }
  for (int idx = threadIdx.x; idx < 8 * 4 * kMlpActions; idx += blockDim.x) {
    const int n = idx % kMlpActions, t = (idx / kMlpActions) % 4, k = idx / (4 * kMlpActions);
    const float* row = w.w3 + n * kMlpHidden + 8 * k + 2 * t;
    s.w3[k][t][n] = n < w.n_actions ? make_float2(to_tf32(__ldg(row)), to_tf32(__ldg(row + 1))) : make_float2(0.0f, 0.0f);
  }
  for (int idx = threadIdx.x; idx < kMlpHidden; idx += blockDim.x) {
    s.b1[idx] = __ldg(w.b1 + idx);
    s.b2[idx] = __ldg(w.b2 + idx);
    if (idx < kMlpActions) s.b3[idx] = idx < w.n_actions ? __ldg(w.b3 + idx) : absent_bias;  // (Q: absent actions never win)
  }
}

// ---- the same network with bf16 operands (fp32 accumulate): mma.sync m16n8k16 --------------------------------------
// Half the MMAs and half the fragment loads of the TF32 form; 8 mantissa bits, so Q-values agree with fp32 only to
// about 1e-2 of their scale: an option for rollouts that tolerate it (S2DMlpPolicy.precision = 1), never the default.
//   A (16x16, row): a0 {(g, 2t), (g, 2t+1)}  a1 {(g+8, ..)}  a2 {(g, 2t+8), (g, 2t+9)}  a3 {(g+8, ..)}   (bf16 pairs)
//   B (16x8, col):  b0 {(k = 2t, n = g), (2t+1, g)}  b1 {(2t+8, g), (2t+9, g)};  D as above
// so two consecutive accumulator column tiles of one layer ARE the A fragment of the next layer's k-step, in order.
// The weight fragments live in the first half of the float2 slots of MlpShared (w1[0], w2[0..3], w3[0..3]).
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], float2 b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)));
}

__device__ __forceinline__ void mlp_load_weights_bf16(MlpShared& s, const MlpWeights& w, float absent_bias) {
  auto w1at = [&](int n, int f) { return f < w.obs_dim ? __ldg(w.w1 + n * w.obs_dim + f) : 0.0f; };
  for (int idx = threadIdx.x; idx < 4 * kMlpHidden; idx += blockDim.x) {
    const int n = idx % kMlpHidden, t = idx / kMlpHidden;
    s.w1[0][t][n] = make_float2(__uint_as_float(pack_bf16(w1at(n, 2 * t), w1at(n, 2 * t + 1))),
                                __uint_as_float(pack_bf16(w1at(n, 2 * t + 8), w1at(n, 2 * t + 9))));
  }
  for (int idx = threadIdx.x; idx < 4 * 4 * kMlpHidden; idx += blockDim.x) {
    const int n = idx % kMlpHidden, t = (idx / kMlpHidden) % 4, k = idx / (4 * kMlpHidden);
    const float* row = w.w2 + n * kMlpHidden + 16 * k + 2 * t;
    s.w2[k][t][n] = make_float2(__uint_as_float(pack_bf16(__ldg(row), __ldg(row + 1))),
                                __uint_as_float(pack_bf16(__ldg(row + 8), __ldg(row + 9))));
  }
  for (int idx = threadIdx.x; idx < 4 * 4 * kMlpActions; idx += blockDim.x) {
    const int n = idx % kMlpActions, t = (idx / kMlpActions) % 4, k = idx / (4 * kMlpActions);
    const float* row = w.w3 + n * kMlpHidden + 16 * k + 2 * t;
    s.w3[k][t][n] = n < w.n_actions ? make_float2(__uint_as_float(pack_bf16(__ldg(row), __ldg(row + 1))),
                                                  __uint_as_float(pack_bf16(__ldg(row + 8), __ldg(row + 9))))
                                    : make_float2(0.0f, 0.0f);
  }
  for (int idx = threadIdx.x; idx < kMlpHidden; idx += blockDim.x) {
    s.b1[idx] = __ldg(w.b1 + idx);
    s.b2[idx] = __ldg(w.b2 + idx);
    if (idx < kMlpActions) s.b3[idx] = idx < w.n_actions ? __ldg(w.b3 + idx) : absent_bias;
  }
}

template <int NT>
__device__ __forceinline__ void mlp_forward_tile_bf16(const MlpShared& s, const float (*obs)[20], int tile, int g, int t,
                                                      float (&q)[NT][4]) {
  float h1[8][4], h2[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h1[j][0] = h1[j][2] = s.b1[8 * j + 2 * t];
    h1[j][1] = h1[j][3] = s.b1[8 * j + 2 * t + 1];
    h2[j][0] = h2[j][2] = s.b2[8 * j + 2 * t];
    h2[j][1] = h2[j][3] = s.b2[8 * j + 2 * t + 1];
  }
  {
    const float* r0 = obs[16 * tile + g] + 2 * t;
    const float* r1 = obs[16 * tile + g + 8] + 2 * t;
    const uint32_t a[4] = {pack_bf16(r0[0], r0[1]), pack_bf16(r1[0], r1[1]), pack_bf16(r0[8], r0[9]), pack_bf16(r1[8], r1[9])};
#pragma unroll
    for (int j = 0; j < 8; ++j) mma_bf16(h1[j], a, s.w1[0][t][8 * j + g]);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t a[4] = {pack_bf16(fmaxf(h1[2 * k][0], 0.0f), fmaxf(h1[2 * k][1], 0.0f)),
                           pack_bf16(fmaxf(h1[2 * k][2], 0.0f), fmaxf(h1[2 * k][3], 0.0f)),
                           pack_bf16(fmaxf(h1[2 * k + 1][0], 0.0f), fmaxf(h1[2 * k + 1][1], 0.0f)),
                           pack_bf16(fmaxf(h1[2 * k + 1][2], 0.0f), fmaxf(h1[2 * k + 1][3], 0.0f))};
#pragma unroll
    for (int j = 0; j < 8; ++j) mma_bf16(h2[j], a, s.w2[k][t][8 * j + g]);
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    q[j][0] = q[j][2] = s.b3[8 * j + 2 * t];
    q[j][1] = q[j][3] = s.b3[8 * j + 2 * t + 1];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t a[4] = {pack_bf16(fmaxf(h2[2 * k][0], 0.0f), fmaxf(h2[2 * k][1], 0.0f)),
                           pack_bf16(fmaxf(h2[2 * k][2], 0.0f), fmaxf(h2[2 * k][3], 0.0f)),
                           pack_bf16(fmaxf(h2[2 * k + 1][0], 0.0f), fmaxf(h2[2 * k + 1][1], 0.0f)),
                           pack_bf16(fmaxf(h2[2 * k + 1][2], 0.0f), fmaxf(h2[2 * k + 1][3], 0.0f))};
#pragma unroll
    for (int j = 0; j < NT; ++j) mma_bf16(q[j], a, s.w3[k][t][8 * j + g]);
  }
}

// Q-values of the 16 episodes `tile` (0 | 1) of this warp, from the staged observations; q[j][..] in the accumulator
// layout: actions 8j + 2t, 8j + 2t + 1 of episode 16 tile + g (q[j][0..1]) and of episode 16 tile + g + 8 (q[j][2..3]).
template <int NT>
__device__ __forceinline__ void mlp_forward_tile(const MlpShared& s, const float (*obs)[20], int tile, int g, int t,
                                                 float (&q)[NT][4]) {
  float h1[8][4], h2[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h1[j][0] = h1[j][2] = s.b1[8 * j + 2 * t];
    h1[j][1] = h1[j][3] = s.b1[8 * j + 2 * t + 1];
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float* r0 = obs[16 * tile + g] + 8 * k + t;
    const float* r1 = obs[16 * tile + g + 8] + 8 * k + t;
    const float a[4] = {to_tf32(r0[0]), to_tf32(r1[0]), to_tf32(r0[4]), to_tf32(r1[4])};
#pragma unroll
    for (int j = 0; j < 8; ++j) mma_tf32(h1[j], a, s.w1[k][t][8 * j + g]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    h2[j][0] = h2[j][2] = s.b2[8 * j + 2 * t];
    h2[j][1] = h2[j][3] = s.b2[8 * j + 2 * t + 1];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {  // accumulator (g, 2t) (g, 2t+1) (g+8, 2t) (g+8, 2t+1) -> A (g, .) (g+8, .) (g, .) (g+8, .)
    const float a[4] = {to_tf32(fmaxf(h1[k][0], 0.0f)), to_tf32(fmaxf(h1[k][2], 0.0f)), to_tf32(fmaxf(h1[k][1], 0.0f)),
                        to_tf32(fmaxf(h1[k][3], 0.0f))};
#pragma unroll
    for (int j = 0; j < 8; ++j) mma_tf32(h2[j], a, s.w2[k][t][8 * j + g]);
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    q[j][0] = q[j][2] = s.b3[8 * j + 2 * t];
    q[j][1] = q[j][3] = s.b3[8 * j + 2 * t + 1];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float a[4] = {to_tf32(fmaxf(h2[k][0], 0.0f)), to_tf32(fmaxf(h2[k][2], 0.0f)), to_tf32(fmaxf(h2[k][1], 0.0f)),
                        to_tf32(fmaxf(h2[k][3], 0.0f))};
#pragma unroll
    for (int j = 0; j < NT; ++j) mma_tf32(q[j], a, s.w3[k][t][8 * j + g]);
  }
}

// greedy action of every episode of the warp -> s.act[warp][episode]; optionally the Q-values to q_out [N][8 NT]
template <int NT, bool BF16 = false>
__device__ __forceinline__ void mlp_greedy(MlpShared& s, int warp, int lane, float* __restrict__ q_out, int64_t warp_first,
                                           int64_t n) {
  const unsigned full = 0xffffffffu;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int tile = 0; tile < 2; ++tile) {
    float q[NT][4];
    if (BF16) mlp_forward_tile_bf16<NT>(s, s.obs[warp], tile, g, t, q);
    else mlp_forward_tile<NT>(s, s.obs[warp], tile, g, t, q);
    if (q_out) {
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int64_t e0 = warp_first + 16 * tile + g, e1 = e0 + 8;
        if (e0 < n) *reinterpret_cast<float2*>(q_out + e0 * (8 * NT) + 8 * j + 2 * t) = make_float2(q[j][0], q[j][1]);
        if (e1 < n) *reinterpret_cast<float2*>(q_out + e1 * (8 * NT) + 8 * j + 2 * t) = make_float2(q[j][2], q[j][3]);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {  // rows g and g + 8
      float best = q[0][2 * half];
      int arg = 2 * t;
#pragma unroll
      for (int j = 0; j < NT; ++j) {  // ascending action index: a strict > keeps the lower index on ties
        if (j > 0 && q[j][2 * half] > best) { best = q[j][2 * half]; arg = 8 * j + 2 * t; }
        if (q[j][2 * half + 1] > best) { best = q[j][2 * half + 1]; arg = 8 * j + 2 * t + 1; }
      }
#pragma unroll
      for (int m = 1; m <= 2; m <<= 1) {  // the quad: same g, t = 0..3; ties go to the lower action index
        const float ob = __shfl_xor_sync(full, best, m);
        const int oa = __shfl_xor_sync(full, arg, m);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
      }
      if (t == 0) s.act[warp][16 * tile + g + 8 * half] = static_cast<uint8_t>(arg);
    }
  }
}

// tanh from the kernels' own exponential (s2d_math.cuh): 1 - 2 / (e^(2|x|) + 1), sign restored
__device__ __forceinline__ float tanh_poly(float x) {
  const float ax = fminf(fabsf(x), 20.0f);
  const float t = 1.0f - 2.0f / (exp_poly(2.0f * ax) + 1.0f);
  return copysignf(t, x);
}

// the actor's action (tanh of the last layer's first `dim` <= 4 outputs) of every episode of the warp -> s.actf
__device__ __forceinline__ void mlp_actor(MlpShared& s, int warp, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
  for (int tile = 0; tile < 2; ++tile) {
    float q[1][4];
    mlp_forward_tile<1>(s, s.obs[warp], tile, g, t, q);
    if (t < 2) {  // columns 2t, 2t + 1 of rows g and g + 8
      *reinterpret_cast<float2*>(&s.actf[warp][16 * tile + g][2 * t]) = make_float2(tanh_poly(q[0][0]), tanh_poly(q[0][1]));
      *reinterpret_cast<float2*>(&s.actf[warp][16 * tile + g + 8][2 * t]) = make_float2(tanh_poly(q[0][2]), tanh_poly(q[0][3]));
    }
  }
}

#ifndef S2D_ROLLOUT_MIN_BLOCKS
#define S2D_ROLLOUT_MIN_BLOCKS 4
#endif

// K closed-loop cycles of a one-player scenario with Discrete actions (ReachBall: n <= 16, Shoot: n <= 24): observe,
// Q-network, (epsilon-)greedy action, step.
// ACT = S2D_ACT_DISCRETE: W is a Q-network, `epsilon` the exploration rate.  ACT = S2D_ACT_CONTINUOUS / S2D_ACT_TURNING
// (ReachBall): W is a DDPG actor (last layer -> tanh -> the Box(1) / Box(4) action) and `epsilon` the half-width of a
// uniform exploration noise added to every action component before the clip to [-1, 1].
template <int SCN, int VAR, int ACT = S2D_ACT_DISCRETE, bool BF16 = false>
__global__ void __launch_bounds__(kBlock, S2D_ROLLOUT_MIN_BLOCKS)
    rollout_mlp_kernel(const __grid_constant__ KernelParams P, const int K, const MlpWeights W, const float epsilon,
                       uint8_t* __restrict__ actions_out, float* __restrict__ q_out, const TrajOut T) {
  constexpr int NT = SCN == S2D_SCENARIO_SHOOT ? 3 : 2;
  constexpr bool kActor = ACT != S2D_ACT_DISCRETE;
  constexpr int kActDim = ACT == S2D_ACT_TURNING ? 4 : 1;
  using SP = typename VariantSP<VAR>::type;
  const SP sp(P.cc);
  __shared__ __align__(16) MlpShared s;
  __shared__ __align__(16) float s_stage[kBlock / 32][32 * kObsDim];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (BF16) mlp_load_weights_bf16(s, W, kActor ? 0.0f : -3.0e38f);
  else mlp_load_weights(s, W, kActor ? 0.0f : -3.0e38f);
  const int64_t i = static_cast<int64_t>(blockIdx.x) * kBlock + threadIdx.x;
  const int64_t n = P.num_envs;
  const bool valid = i < n;
  const uint64_t gid = static_cast<uint64_t>(P.env_id_offset + i);
  const int64_t warp_first = i - lane;
  const int64_t il = valid ? i : n - 1;
#pragma unroll
  for (int f = kObsDim; f < 16; ++f) s.obs[warp][lane][f] = 0.0f;
  __syncthreads();
  if (warp_first >= n) return;  // (whole warp; after the only block-wide barrier)

  Episode e;
  LaunchOut out;
  float obs_row[kObsDim];
  load_episode(P.state, n, il, e);
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    scenario_obs<SCN>(e, obs_row);
#pragma unroll
    for (int f = 0; f < kObsDim; ++f) s.obs[warp][lane][f] = obs_row[f];
    __syncwarp();
    if (T.obs) {  // the staged rows of the warp, 320 contiguous floats
      float* dst = T.obs + (static_cast<int64_t>(k) * n + warp_first) * kObsDim;
      const int rows = static_cast<int>(n - warp_first < 32 ? n - warp_first : 32);
#pragma unroll
      for (int m = 0; m < kObsDim; ++m) {
        const int idx = lane + 32 * m, row = idx / kObsDim;
        if (row < rows) dst[idx] = s.obs[warp][row][idx - row * kObsDim];
      }
    }
    int a = 0;
    int rs;
    const float reward_before = out.reward_sum;  // (this cycle's reward is taken out exactly: the sum restarts at 0 ...
    if (!kActor) {
      mlp_greedy<NT, BF16>(s, warp, lane, k == K - 1 ? q_out : nullptr, warp_first, n);
      __syncwarp();
      a = s.act[warp][lane];
      if (epsilon > 0.0f) {  // exploration: the same counter stream as the turning action's draw (RNG_ACTION)
        const uint4 w = philox4x32_10(P.seed, gid, e.cycle, RNG_ACTION, 1);
        if (u32_to_unit(w.x) < epsilon) a = u32_to_int(w.y, 0, W.n_actions - 1);
      }
      if (actions_out && valid) actions_out[i * K + k] = static_cast<uint8_t>(a);
      out.reward_sum = 0.0f;
      if (SCN == S2D_SCENARIO_SHOOT) {
        const float4 tab = __ldg(P.action_table + a);
        rs = substep<SCN, S2D_ACT_DISCRETE, SP, true>(e, P, sp, gid, i, tab.x, tab.y, tab.z, tab.w, out);
      } else {
        const float2 tab = __ldg(reinterpret_cast<const float2*>(P.action_table + a) + 1);
        rs = substep<SCN, S2D_ACT_DISCRETE, SP, true>(e, P, sp, gid, i, 0.f, 0.f, tab.x, tab.y, out, P.sincos_memo);
      }
    } else {
      mlp_actor(s, warp, lane);
      __syncwarp();
      float4 av = *reinterpret_cast<const float4*>(s.actf[warp][lane]);
      if (epsilon > 0.0f) {
        const uint4 w = philox4x32_10(P.seed, gid, e.cycle, RNG_ACTION, 2);
        av.x = clampf(-1.0f, av.x + epsilon * u11(w.x), 1.0f);
        av.y = clampf(-1.0f, av.y + epsilon * u11(w.y), 1.0f);
        av.z = clampf(-1.0f, av.z + epsilon * u11(w.z), 1.0f);
        av.w = clampf(-1.0f, av.w + epsilon * u11(w.w), 1.0f);
      }
      if (T.actions_f && valid) {
        float* dst = T.actions_f + (static_cast<int64_t>(k) * n + i) * kActDim;
        dst[0] = av.x;
        if (kActDim == 4) { dst[1] = av.y; dst[2] = av.z; dst[3] = av.w; }
      }
      out.reward_sum = 0.0f;
      rs = substep<SCN, ACT, SP, true>(e, P, sp, gid, i, av.x, kActDim == 4 ? av.y : 0.f, kActDim == 4 ? av.z : 0.f,
                                       kActDim == 4 ? av.w : 0.f, out, kActDim == 1 ? P.sincos_memo : nullptr);
    }
    const float rw = out.reward_sum;
    out.reward_sum = reward_before + rw;  // ... and is put back together in the order the step kernel adds it up)
    if (valid) {
      const int64_t at = static_cast<int64_t>(k) * n + i;
      if (T.actions) T.actions[at] = static_cast<uint8_t>(a);
      if (T.reward) T.reward[at] = rw;
      if (T.done) T.done[at] = static_cast<uint8_t>(rs != S2D_RESULT_NONE);
    }
    end_of_episode<SCN, SP, false, S2D_ROLLOUT_DRAW_MEMO>(e, P, sp, gid, i, valid, rs, out);
  }
  if (valid) {
    store_episode(P.state, n, i, e);
    scenario_obs<SCN>(e, obs_row);
    if (T.obs) {
      float* dst = T.obs + (static_cast<int64_t>(K) * n + i) * kObsDim;
#pragma unroll
      for (int f = 0; f < kObsDim; ++f) dst[f] = obs_row[f];
    }
    P.reward[i] = out.reward_sum;
    P.done[i] = static_cast<uint8_t>(out.ended != 0);
    P.result[i] = static_cast<uint8_t>(out.last_result());
  } else {
    out = LaunchOut();
  }
  warp_store_obs(P.obs, warp_first, n, obs_row, valid, s_stage[warp]);
  flush_tally(out, P.stats);
}

}  // namespace s2d
