// emu_reachball.cpp - TEST-ONLY: runs the __device__ functions of csrc/s2d_scenarios.cuh (substep, reset_episode,
// build_obs, load/store_episode) on the host, one "thread" after another, over a state buffer with the same
// plane-major layout as the GPU's.  Lets the CPU test suite check the kernel source against the fp32 oracle.
#define S2D_HOST_EMU 1
#include "cuda_shim.h"
#include "../../gym-soccer-2d-env_b200/csrc/s2d_scenarios.cuh"

#include <vector>

using namespace s2d;

struct Emu {
  KernelParams kp;
  float4 table[256];
  float2 memo[kSinCosMemoSize];  // sincos_deg over the whole degrees, as fill_sincos_memo (s2d_api.cu) writes it
  bool default_sp;  // same dispatch as s2d_create: constant-folded accessors for the default ServerParam
  std::vector<unsigned char> state;
};

template <int SCN, int ACT, class SP>
static void step_all(Emu* h, const void* actions, int K, float* obs, float* reward, uint8_t* done, uint8_t* result,
                     double* stats) {
  const KernelParams& P = h->kp;
  const SP sp(P.cc);
  const int64_t n = P.num_envs;
  for (int64_t i = 0; i < n; ++i) {
    Episode e;
    LaunchOut out;
    load_episode(P.state, n, i, e);
    const uint64_t gid = (uint64_t)(P.env_id_offset + i);
    // step_kernel's rule: ReachBall Discrete(n) launches of K >= S2D_SINCOS_MEMO_MIN_K (4) cycles read the memo
    const float2* memo = (ACT == S2D_ACT_DISCRETE && SCN != S2D_SCENARIO_SHOOT && K >= 4) ? h->memo : nullptr;
    for (int k = 0; k < K; ++k) {
      if (ACT == S2D_ACT_DISCRETE) {
        const float4 t = h->table[((const uint8_t*)actions)[i * K + k]];
        substep<SCN, ACT>(e, P, sp, gid, i, t.x, t.y, t.z, t.w, out, memo);
      } else if (ACT == S2D_ACT_CONTINUOUS) {
        substep<SCN, ACT>(e, P, sp, gid, i, ((const float*)actions)[i * K + k], 0.f, 0.f, 0.f, out);
      } else {
        const float* a = (const float*)actions + (i * K + k) * 4;
        substep<SCN, ACT>(e, P, sp, gid, i, a[0], a[1], a[2], a[3], out);
      }
    }
    store_episode(P.state, n, i, e);
    scenario_obs<SCN>(e, obs + i * kObsDim);
    reward[i] = out.reward_sum;
    done[i] = (uint8_t)(out.ended != 0);
    result[i] = (uint8_t)out.last_result();
    stats[0] += out.episodes(); stats[1] += out.goals(); stats[2] += out.outs(); stats[3] += out.timeouts();
    stats[4] += out.ep_steps; stats[5] += out.ret;
  }
}

extern "C" {

void* emu_create(const S2DConfig* cfg) {
  Emu* h = new Emu();
  make_kernel_params(*cfg, h->kp, h->table);
  h->default_sp = is_default_server_param(cfg->sp) && cfg->collision_model == S2D_COLLISION_MIDPOINT;
  h->state.assign((size_t)cfg->num_envs * kStateBytesPerEnv, 0);
  h->kp.state = h->state.data();
  h->kp.action_table = h->table;
  for (int j = 0; j < kSinCosMemoSize; ++j) sincos_deg((float)(j - kSinCosMemoHalf), h->memo[j].x, h->memo[j].y);
  h->kp.sincos_memo = h->memo;
  return h;
}
void emu_destroy(void* p) { delete (Emu*)p; }
void* emu_state(void* p) { return ((Emu*)p)->state.data(); }

void emu_reset(void* p, const uint8_t* mask, float* obs) {
  Emu* h = (Emu*)p;
  const KernelParams& P = h->kp;
  for (int64_t i = 0; i < P.num_envs; ++i) {
    if (mask && !mask[i]) continue;
    Episode e;
    load_episode(P.state, P.num_envs, i, e);
    const uint64_t gid = (uint64_t)(P.env_id_offset + i);
    if (P.scenario == S2D_SCENARIO_SHOOT) {
      if (P.noise) reset_episode<S2D_SCENARIO_SHOOT>(e, P, NoisySP(P.cc), gid);
      else reset_episode<S2D_SCENARIO_SHOOT>(e, P, RuntimeSP(P.cc), gid);
      scenario_obs<S2D_SCENARIO_SHOOT>(e, obs + i * kObsDim);
    } else {
      if (P.noise) reset_episode<S2D_SCENARIO_REACHBALL>(e, P, NoisySP(P.cc), gid);
      else reset_episode<S2D_SCENARIO_REACHBALL>(e, P, RuntimeSP(P.cc), gid);
      scenario_obs<S2D_SCENARIO_REACHBALL>(e, obs + i * kObsDim);
    }
    store_episode(P.state, P.num_envs, i, e);
  }
}

void emu_step(void* p, const void* actions, int K, float* obs, float* reward, uint8_t* done, uint8_t* result,
              float* terminal_obs, double* stats6) {
  Emu* h = (Emu*)p;
  h->kp.terminal_obs = terminal_obs;
#define EMU_STEP(SCN, ACT)                                                                            \
  do {                                                                                                \
    if (h->kp.noise) step_all<SCN, ACT, NoisySP>(h, actions, K, obs, reward, done, result, stats6);     \
    else if (h->default_sp) step_all<SCN, ACT, DefaultSP>(h, actions, K, obs, reward, done, result, stats6); \
    else step_all<SCN, ACT, RuntimeSP>(h, actions, K, obs, reward, done, result, stats6);             \
  } while (0)
  if (h->kp.scenario == S2D_SCENARIO_SHOOT) {
    if (h->kp.action_mode == S2D_ACT_DISCRETE) EMU_STEP(S2D_SCENARIO_SHOOT, S2D_ACT_DISCRETE);
    else EMU_STEP(S2D_SCENARIO_SHOOT, S2D_ACT_COMMAND);
  } else {
    switch (h->kp.action_mode) {
      case S2D_ACT_DISCRETE: EMU_STEP(S2D_SCENARIO_REACHBALL, S2D_ACT_DISCRETE); break;
      case S2D_ACT_CONTINUOUS: EMU_STEP(S2D_SCENARIO_REACHBALL, S2D_ACT_CONTINUOUS); break;
      case S2D_ACT_TURNING: EMU_STEP(S2D_SCENARIO_REACHBALL, S2D_ACT_TURNING); break;
      default: EMU_STEP(S2D_SCENARIO_REACHBALL, S2D_ACT_COMMAND); break;
    }
  }
}

// DefaultSP (compile-time constants) must equal make_cycle_consts(defaults) bit for bit; returns the mismatches
int emu_check_default_consts(void) {
  S2DServerParam sp;
  default_server_param(sp);
  const CycleConsts c = make_cycle_consts(sp);
  const RuntimeSP r(c);
  const DefaultSP d(c);
  int bad = 0;
#define X(name, def) bad += memcmp(&(const float&)r.name(), &(const float&)d.name(), 4) != 0;
#define CHK(name) { float a = r.name(), b = d.name(); bad += memcmp(&a, &b, 4) != 0; }
#undef X
#define X(name, def) CHK(name)
  S2D_SERVER_PARAMS(X)
#undef X
#define Y(name, expr) CHK(name)
  S2D_DERIVED_PARAMS(Y)
#undef Y
  return bad;
}

int emu_uses_default_sp(void* p) { return ((Emu*)p)->default_sp ? 1 : 0; }

}  // extern "C"
