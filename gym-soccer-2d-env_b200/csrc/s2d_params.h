// s2d_params.h - the physics constants of the cycle, in the two forms the kernels use them.
//
// Names = proto ServerParam / PlayerType fields (idl/service.proto:1435-1732); the values are NOT in the
// reference tree (they arrive from rcssserver at run time and are stored unused, server.py:105-118), so the
// defaults below are rcssserver's documented ones.
//
//   RuntimeSP   reads S2DServerParam (+ derived products) from the kernel's constant bank: any configuration
//   NoisySP     RuntimeSP, and the physics functions add noise (S2DConfig.noise != 0)
//   DefaultSP   the same accessors as compile-time constants for the default configuration: immediates, no
//               constant loads, dead branches folded (back-dash, slowness, angle step ...).  s2d_create picks
//               the DefaultSP kernels when cfg->sp equals s2d_default_server_param() bit for bit.
// Both give bit-identical results for the default configuration (checked in tests/emu and on the GPU).
#pragma once
#include "../../include/soccer2d.h"

// X(name, rcssserver default)
#define S2D_SERVER_PARAMS(X)                                                                                  \
  X(pitch_half_length, 52.5f) X(pitch_half_width, 34.0f) X(goal_width, 14.02f) X(goal_post_radius, 0.06f)     \
  X(ball_size, 0.085f) X(ball_decay, 0.94f) X(ball_rand, 0.05f) X(ball_speed_max, 3.0f) X(ball_accel_max, 2.7f) \
  X(player_size, 0.3f) X(player_decay, 0.4f) X(player_rand, 0.1f) X(player_speed_max, 1.05f)                  \
  X(player_accel_max, 1.0f) X(dash_power_rate, 0.006f) X(inertia_moment, 5.0f)                                \
  X(min_dash_power, 0.0f) X(max_dash_power, 100.0f) X(min_dash_angle, -180.0f) X(max_dash_angle, 180.0f)      \
  X(dash_angle_step, 1.0f) X(side_dash_rate, 0.4f) X(back_dash_rate, 0.7f)                                    \
  X(min_power, -100.0f) X(max_power, 100.0f) X(min_moment, -180.0f) X(max_moment, 180.0f)                     \
  X(kick_power_rate, 0.027f) X(kickable_margin, 0.7f) X(kick_rand, 0.1f)                                      \
  X(stamina_max, 8000.0f) X(stamina_inc_max, 45.0f) X(extra_stamina, 50.0f) X(stamina_capacity, 130600.0f)    \
  X(recover_init, 1.0f) X(recover_min, 0.5f) X(recover_dec, 0.002f) X(recover_dec_thr, 0.3f)                  \
  X(effort_init, 1.0f) X(effort_max, 1.0f) X(effort_min, 0.6f) X(effort_dec, 0.005f) X(effort_dec_thr, 0.3f)  \
  X(effort_inc, 0.01f) X(effort_inc_thr, 0.6f)                                                                \
  X(slowness_on_top_for_left_team, 1.0f) X(slowness_on_top_for_right_team, 1.0f)

// derived constants: Y(name, expression over the accessors p.x())
#define S2D_DERIVED_PARAMS(Y)                                                                                 \
  Y(inv_dash_angle_step, p.dash_angle_step() > 1.0e-10f ? static_cast<float>(1.0 / static_cast<double>(p.dash_angle_step())) : 0.0f) \
  Y(player_accel_max2, p.player_accel_max() * p.player_accel_max())                                           \
  Y(player_speed_max2, p.player_speed_max() * p.player_speed_max())                                           \
  Y(ball_accel_max2, p.ball_accel_max() * p.ball_accel_max())                                                 \
  Y(ball_speed_max2, p.ball_speed_max() * p.ball_speed_max())                                                 \
  Y(collide_r, p.player_size() + p.ball_size())                                                               \
  Y(collide_r2, (p.player_size() + p.ball_size()) * (p.player_size() + p.ball_size()))                        \
  Y(recover_dec_stamina, p.recover_dec_thr() * p.stamina_max())                                               \
  Y(effort_dec_stamina, p.effort_dec_thr() * p.stamina_max())                                                 \
  Y(effort_inc_stamina, p.effort_inc_thr() * p.stamina_max())                                                 \
  Y(kickable_area, p.player_size() + p.ball_size() + p.kickable_margin())

#if defined(__CUDACC__)
#define S2D_HD __host__ __device__ __forceinline__
#define S2D_HDC __host__ __device__ static constexpr
#elif defined(S2D_HOST_EMU)
#define S2D_HD inline __attribute__((always_inline))
#define S2D_HDC static constexpr
#else
#define S2D_HD inline
#define S2D_HDC static constexpr
#endif

namespace s2d {

// raw accessors over an S2DServerParam in memory (host: to evaluate the derived expressions)
struct RawSP {
  const S2DServerParam& s;
#define X(name, def) S2D_HD float name() const { return s.name; }
  S2D_SERVER_PARAMS(X)
#undef X
};

// What a kernel receives: the proto-named values plus the derived products, computed once on the host in
// float arithmetic (so device and host agree on every bit).
struct CycleConsts {
  S2DServerParam sp;
  int collision_model;  // S2D_COLLISION_* (set by make_kernel_params)
#define Y(name, expr) float name;
  S2D_DERIVED_PARAMS(Y)
#undef Y
};

inline CycleConsts make_cycle_consts(const S2DServerParam& sp) {
  CycleConsts c;
  c.sp = sp;
  c.collision_model = S2D_COLLISION_MIDPOINT;
  const RawSP p{sp};
#define Y(name, expr) c.name = (expr);
  S2D_DERIVED_PARAMS(Y)
#undef Y
  return c;
}

inline void default_server_param(S2DServerParam& sp) {
  sp = S2DServerParam{};
#define X(name, def) sp.name = def;
  S2D_SERVER_PARAMS(X)
#undef X
}

inline bool is_default_server_param(const S2DServerParam& sp) {
  bool same = true;
#define X(name, def) same = same && (sp.name == def);
  S2D_SERVER_PARAMS(X)
#undef X
  return same;
}

// accessors over the kernel's constant bank
struct RuntimeSP {
  static constexpr bool kNoise = false;
  static constexpr bool kHetero = false;
  const CycleConsts& c;
  S2D_HD explicit RuntimeSP(const CycleConsts& c_) : c(c_) {}
  S2D_HD int collision_model() const { return c.collision_model; }
#define X(name, def) S2D_HD float name() const { return c.sp.name; }
  S2D_SERVER_PARAMS(X)
#undef X
#define Y(name, expr) S2D_HD float name() const { return c.name; }
  S2D_DERIVED_PARAMS(Y)
#undef Y
};

// same constants, and the cycle adds rcssserver's noise (player_rand, ball_rand, kick_rand) from the counter RNG
struct NoisySP : RuntimeSP {
  static constexpr bool kNoise = true;
  S2D_HD explicit NoisySP(const CycleConsts& c_) : RuntimeSP(c_) {}
};

// Heterogeneous players (FULLGAME): the per-player fields of S2DPlayerType come from this lane's row of the type table
// (shared memory, 16 floats a row: S2DPlayerType with the derived kickable_area in reserved[0]); the rest as RuntimeSP.
enum { PT_PLAYER_DECAY, PT_INERTIA_MOMENT, PT_DASH_POWER_RATE, PT_STAMINA_INC_MAX, PT_KICKABLE_MARGIN, PT_KICK_RAND,
       PT_EXTRA_STAMINA, PT_EFFORT_MAX, PT_EFFORT_MIN, PT_KICK_POWER_RATE, PT_KICKABLE_AREA, PT_ROW = 16 };
template <bool NOISE>
struct HeteroSP : RuntimeSP {
  static constexpr bool kNoise = NOISE;
  static constexpr bool kHetero = true;
  const float* row = nullptr;    // this player's type
  const float* table = nullptr;  // all types: [S2D_MAX_PLAYER_TYPES][PT_ROW]
  S2D_HD explicit HeteroSP(const CycleConsts& c_) : RuntimeSP(c_) {}
  S2D_HD float player_decay() const { return row[PT_PLAYER_DECAY]; }
  S2D_HD float inertia_moment() const { return row[PT_INERTIA_MOMENT]; }
  S2D_HD float dash_power_rate() const { return row[PT_DASH_POWER_RATE]; }
  S2D_HD float stamina_inc_max() const { return row[PT_STAMINA_INC_MAX]; }
  S2D_HD float kickable_margin() const { return row[PT_KICKABLE_MARGIN]; }
  S2D_HD float kick_rand() const { return row[PT_KICK_RAND]; }
  S2D_HD float extra_stamina() const { return row[PT_EXTRA_STAMINA]; }
  S2D_HD float effort_max() const { return row[PT_EFFORT_MAX]; }
  S2D_HD float effort_init() const { return row[PT_EFFORT_MAX]; }
  S2D_HD float effort_min() const { return row[PT_EFFORT_MIN]; }
  S2D_HD float kick_power_rate() const { return row[PT_KICK_POWER_RATE]; }
  S2D_HD float kickable_area() const { return row[PT_KICKABLE_AREA]; }
};

// the default configuration as compile-time constants
struct DefaultBase {
#define X(name, def) S2D_HDC float name() { return def; }
  S2D_SERVER_PARAMS(X)
#undef X
};
struct DefaultSP : DefaultBase {
  S2D_HDC int collision_model() { return S2D_COLLISION_MIDPOINT; }  // (other models run the RuntimeSP kernels)
  static constexpr bool kNoise = false;
  static constexpr bool kHetero = false;
  S2D_HD explicit DefaultSP(const CycleConsts&) {}
#define Y(name, expr)                \
  S2D_HDC float name() {             \
    constexpr DefaultBase p{};       \
    return (expr);                   \
  }
  S2D_DERIVED_PARAMS(Y)
#undef Y
};

}  // namespace s2d
