// emu_reachball.cpp - TEST-ONLY: runs the __device__ functions of csrc/s2d_reachball.cuh (substep, reset_episode,
// build_obs, load/store_episode) on the host, one "thread" after another, over a state buffer with the same
// plane-major layout as the GPU's.  Lets the CPU test suite check the kernel source against the fp32 oracle.
#define S2D_HOST_EMU 1
#include "cuda_shim.h"
#include "../../gym-soccer-2d-env_b200/csrc/s2d_reachball.cuh"

#include <vector>

using namespace s2d;

struct Emu {
  KernelParams kp;
  float dirs[256];
  std::vector<unsigned char> state;
};

template <int ACT>
static void step_all(Emu* h, const void* actions, int K, float* obs, float* reward, uint8_t* done, uint8_t* result,
                     double* stats) {
  const KernelParams& P = h->kp;
  const int64_t n = P.num_envs;
  for (int64_t i = 0; i < n; ++i) {
    Episode e;
    LaunchOut out;
    load_episode(P.state, n, i, e);
    const uint64_t gid = (uint64_t)(P.env_id_offset + i);
    for (int k = 0; k < K; ++k) {
      if (ACT == S2D_ACT_DISCRETE) {
        substep<ACT>(e, P, gid, i, h->dirs[((const uint8_t*)actions)[i * K + k]], 0.f, 0.f, 0.f, out);
      } else if (ACT == S2D_ACT_CONTINUOUS) {
        substep<ACT>(e, P, gid, i, ((const float*)actions)[i * K + k], 0.f, 0.f, 0.f, out);
      } else {
        const float* a = (const float*)actions + (i * K + k) * 4;
        substep<ACT>(e, P, gid, i, a[0], a[1], a[2], a[3], out);
      }
    }
    store_episode(P.state, n, i, e);
    build_obs(e, obs + i * kObsDim);
    reward[i] = out.reward_sum;
    done[i] = (uint8_t)out.any_done;
    result[i] = (uint8_t)out.last_result;
    stats[0] += out.episodes; stats[1] += out.goals; stats[2] += out.outs; stats[3] += out.timeouts;
    stats[4] += out.ep_steps; stats[5] += out.ret;
  }
}

extern "C" {

void* emu_create(const S2DConfig* cfg) {
  Emu* h = new Emu();
  make_kernel_params(*cfg, h->kp, h->dirs);
  h->state.assign((size_t)cfg->num_envs * kStateBytesPerEnv, 0);
  h->kp.state = h->state.data();
  h->kp.dash_dirs = h->dirs;
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
    reset_episode(e, P, (uint64_t)(P.env_id_offset + i));
    store_episode(P.state, P.num_envs, i, e);
    build_obs(e, obs + i * kObsDim);
  }
}

void emu_step(void* p, const void* actions, int K, float* obs, float* reward, uint8_t* done, uint8_t* result,
              float* terminal_obs, double* stats6) {
  Emu* h = (Emu*)p;
  h->kp.terminal_obs = terminal_obs;
  switch (h->kp.action_mode) {
    case S2D_ACT_DISCRETE: step_all<S2D_ACT_DISCRETE>(h, actions, K, obs, reward, done, result, stats6); break;
    case S2D_ACT_CONTINUOUS: step_all<S2D_ACT_CONTINUOUS>(h, actions, K, obs, reward, done, result, stats6); break;
    default: step_all<S2D_ACT_TURNING>(h, actions, K, obs, reward, done, result, stats6); break;
  }
}

}  // extern "C"
