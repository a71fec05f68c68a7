// s2d_one_player.cuh - per-episode state and the rcssserver-style cycle for the one-player scenarios
// (ReachBall: sample_environments/reach_ball_env.py; Shoot: 1v0 with the kick model).
//
// One thread owns one episode; everything below lives in registers between the state load and the state
// store of a launch.  The arithmetic follows the fp32 SPEC of s2d_math.cuh (no implicit fma).
//
// What each function stands in for (the reference only *launches* these; soccer_2d_env.py:356-398):
//   dash / turn / kick        rcssserver Player::dash, Player::turn, Player::kick
//   move_object               rcssserver MPObject::_inc (accel clamp, speed clamp, move, decay)
//   collide_ball_player       rcssserver Stadium::collisions specialised to one player and the ball
//   update_stamina            rcssserver Player::updateStamina
//   recover                   trainer (recover) = proto DoRecover (idl/service.proto:1407)
#pragma once
#include "../../include/soccer2d.h"
#include "s2d_math.cuh"

namespace s2d {

// HBM layout: five 16-byte planes, plane-major (plane p of env i at state + (p * N + i) * 16), so that a
// warp reads / writes 512 contiguous bytes per plane with one 128-bit access per lane.
//   plane 0  float4 {player x, y, vx, vy}
//   plane 1  float4 {body_direction (deg), stamina, effort, recovery}
//   plane 2  float4 {ball x, y, vx, vy}
//   plane 3  float4 {distance_to_ball memory, body_ball_angle_diff memory, stamina_capacity, episode return}
//   plane 4  uint4  {step_number, cycle, episode, flags | play_mode << 8 | scores << 16}
constexpr int kStatePlanes = 5;
constexpr int kStateBytesPerEnv = kStatePlanes * 16;

struct Episode {
  float px, py, vx, vy;
  float body, stamina, effort, recovery;
  float bx, by, bvx, bvy;
  float mem_dist, mem_ang, capacity, ep_return;
  int step_number;
  uint32_t cycle, episode, flags;
};

__device__ __forceinline__ void load_episode(const void* state, int64_t n, int64_t i, Episode& e) {
  const float4* f = reinterpret_cast<const float4*>(state);
  const float4 a = ld_stream(f + i), b = ld_stream(f + n + i), c = ld_stream(f + 2 * n + i),
               d = ld_stream(f + 3 * n + i);
  const uint4 u = ld_stream(reinterpret_cast<const uint4*>(state) + 4 * n + i);
  e.px = a.x; e.py = a.y; e.vx = a.z; e.vy = a.w;
  e.body = b.x; e.stamina = b.y; e.effort = b.z; e.recovery = b.w;
  e.bx = c.x; e.by = c.y; e.bvx = c.z; e.bvy = c.w;
  e.mem_dist = d.x; e.mem_ang = d.y; e.capacity = d.z; e.ep_return = d.w;
  e.step_number = static_cast<int>(u.x); e.cycle = u.y; e.episode = u.z; e.flags = u.w;
}

__device__ __forceinline__ void store_episode(void* state, int64_t n, int64_t i, const Episode& e) {
  float4* f = reinterpret_cast<float4*>(state);
  st_stream(f + i, make_float4(e.px, e.py, e.vx, e.vy));
  st_stream(f + n + i, make_float4(e.body, e.stamina, e.effort, e.recovery));
  st_stream(f + 2 * n + i, make_float4(e.bx, e.by, e.bvx, e.bvy));
  st_stream(f + 3 * n + i, make_float4(e.mem_dist, e.mem_ang, e.capacity, e.ep_return));
  st_stream(reinterpret_cast<uint4*>(state) + 4 * n + i,
            make_uint4(static_cast<uint32_t>(e.step_number), e.cycle, e.episode, e.flags));
}

// ---- commands -------------------------------------------------------------------------------------

__device__ __forceinline__ void recover(Episode& e, const S2DServerParam& sp) {
  e.stamina = sp.stamina_max;
  e.recovery = sp.recover_init;
  e.effort = sp.effort_max;
  e.capacity = sp.stamina_capacity;
}

// (dash power dir): stamina is charged first, then the effective power is scaled by effort, the
// direction-dependent rate (forward 1, sideways side_dash_rate, backwards back_dash_rate) and
// dash_power_rate.  Returns the acceleration it adds.
__device__ __forceinline__ void dash(Episode& e, float power, float dir, const S2DServerParam& sp, float& ax,
                                     float& ay) {
  power = clampf(sp.min_dash_power, power, sp.max_dash_power);
  dir = clampf(sp.min_dash_angle, dir, sp.max_dash_angle);
  if (sp.dash_angle_step > 1.0e-10f) dir = sp.dash_angle_step * rintf(dir / sp.dash_angle_step);
  const bool back = power < 0.0f;
  float need = back ? power * -2.0f : power;
  need = fmin_(need, e.stamina + sp.extra_stamina);
  e.stamina = fmax_(0.0f, e.stamina - need);
  power = back ? need / -2.0f : need;
  const float ad = fabsf(dir);
  float rate;
  if (ad > 90.0f)
    rate = sp.back_dash_rate - ((sp.back_dash_rate - sp.side_dash_rate) * (1.0f - (ad - 90.0f) / 90.0f));
  else
    rate = sp.side_dash_rate + ((1.0f - sp.side_dash_rate) * (1.0f - ad / 90.0f));
  rate = clampf(0.0f, rate, 1.0f);
  float eff = fabsf(e.effort * power * rate * sp.dash_power_rate);
  if (e.py < 0.0f) {
    const float slow = sp.slowness_on_top_for_left_team;  // the single player is on the left team
    if (slow != 1.0f) eff /= slow;
  }
  if (back) dir += 180.0f;
  float s, c;
  sincos_deg(e.body + dir, s, c);
  ax += eff * c;
  ay += eff * s;
}

// (turn moment): the faster the player moves, the less it turns (inertia_moment)
__device__ __forceinline__ void turn(Episode& e, float moment, const S2DServerParam& sp) {
  moment = clampf(sp.min_moment, moment, sp.max_moment);
  const float speed = hypot2(e.vx, e.vy);
  e.body = norm_deg(e.body + moment / (1.0f + sp.inertia_moment * speed));
}

// (kick power dir): only inside the kickable area; power falls off by up to 25 % with the angle between
// body and ball and by up to 25 % with the distance.  Adds to the ball's acceleration.
__device__ __forceinline__ bool kick(Episode& e, float power, float dir, const S2DServerParam& sp, float& bax,
                                     float& bay) {
  power = clampf(0.0f, power, sp.max_power);
  dir = clampf(sp.min_moment, dir, sp.max_moment);
  const float dx = e.bx - e.px, dy = e.by - e.py;
  const float dist = hypot2(dx, dy);
  if (dist > sp.player_size + sp.ball_size + sp.kickable_margin) return false;
  const float dir_diff = fabsf(norm_deg(atan2_deg(dy, dx) - e.body));
  const float dist_ball = dist - sp.player_size - sp.ball_size;
  const float eff =
      power * sp.kick_power_rate * (1.0f - 0.25f * dir_diff / 180.0f - 0.25f * dist_ball / sp.kickable_margin);
  float s, c;
  sincos_deg(e.body + dir, s, c);
  bax += eff * c;
  bay += eff * s;
  return true;
}

// ---- one cycle ------------------------------------------------------------------------------------

__device__ __forceinline__ void move_object(float& x, float& y, float& vx, float& vy, float ax, float ay,
                                            float accel_max, float speed_max, float decay) {
  if (ax != 0.0f || ay != 0.0f) {
    const float a = hypot2(ax, ay);
    if (a > accel_max) {
      const float k = accel_max / a;
      ax *= k;
      ay *= k;
    }
    vx += ax;
    vy += ay;
    const float v = hypot2(vx, vy);
    if (v > speed_max) {
      const float k = speed_max / v;
      vx *= k;
      vy *= k;
    }
  }
  x += vx;
  y += vy;
  vx *= decay;
  vy *= decay;
}

constexpr float kCollideEps = 1.0e-6f;

// Ball overlapping the player: the ball goes back along its own velocity until the two just touch; if it
// is not moving (or the line misses), it is pushed out radially.  The player keeps its place.  Up to ten
// relaxation rounds as in the server, then both objects that collided get vel *= -0.1.
__device__ __forceinline__ void collide_ball_player(Episode& e, const S2DServerParam& sp) {
  uint32_t hit = 0;
  const float r = sp.player_size + sp.ball_size;
#pragma unroll 1
  for (int round = 0; round < 10; ++round) {
    const float dx = e.bx - e.px, dy = e.by - e.py;
    if (!(dx * dx + dy * dy < r * r)) break;
    hit = S2D_FLAG_BALL_COLLIDED | S2D_FLAG_PLAYER_COLLIDED;
    const float rr = r + kCollideEps;
    const float v = hypot2(e.bvx, e.bvy);
    bool placed = false;
    if (v > 1.0e-10f) {
      const float ux = e.bvx / v, uy = e.bvy / v;
      const float du = dx * ux + dy * uy;
      const float disc = du * du - (dx * dx + dy * dy - rr * rr);
      if (disc >= 0.0f) {
        const float t = du + sqrtf(disc);
        if (t >= 0.0f) {
          e.bx = e.bx - t * ux;
          e.by = e.by - t * uy;
          placed = true;
        }
      }
    }
    if (!placed) {
      const float d = hypot2(dx, dy);
      if (d < 1.0e-10f) {
        e.bx = e.px + rr;
        e.by = e.py;
      } else {
        e.bx = e.px + dx / d * rr;
        e.by = e.py + dy / d * rr;
      }
    }
  }
  if (hit) {
    e.bvx *= -0.1f;
    e.bvy *= -0.1f;
    e.vx *= -0.1f;
    e.vy *= -0.1f;
  }
  e.flags = (e.flags & ~(S2D_FLAG_BALL_COLLIDED | S2D_FLAG_PLAYER_COLLIDED)) | hit;
}

__device__ __forceinline__ void update_stamina(Episode& e, const S2DServerParam& sp) {
  if (e.stamina <= sp.recover_dec_thr * sp.stamina_max) {
    if (e.recovery > sp.recover_min) e.recovery -= sp.recover_dec;
    if (e.recovery < sp.recover_min) e.recovery = sp.recover_min;
  }
  if (e.stamina <= sp.effort_dec_thr * sp.stamina_max) {
    if (e.effort > sp.effort_min) e.effort -= sp.effort_dec;
    if (e.effort < sp.effort_min) e.effort = sp.effort_min;
  }
  if (e.stamina >= sp.effort_inc_thr * sp.stamina_max) {
    if (e.effort < sp.effort_max) {
      e.effort += sp.effort_inc;
      if (e.effort > sp.effort_max) e.effort = sp.effort_max;
    }
  }
  float inc = fmin_(e.recovery * sp.stamina_inc_max, sp.stamina_max - e.stamina);
  if (sp.stamina_capacity >= 0.0f) {
    if (inc > e.capacity) inc = e.capacity;
  }
  e.stamina += inc;
  if (sp.stamina_capacity >= 0.0f) e.capacity = fmax_(0.0f, e.capacity - inc);
}

// One server cycle with at most one body command.  Order as in rcssserver's Stadium::step: commands were
// applied on receipt, then every object moves, then collisions, then stamina, then the clock.
__device__ __forceinline__ void simulate_cycle(Episode& e, int cmd, float power, float dir,
                                               const S2DServerParam& sp) {
  float ax = 0.0f, ay = 0.0f, bax = 0.0f, bay = 0.0f;
  e.flags &= ~S2D_FLAG_KICKED;
  if (cmd == S2D_CMD_DASH) {
    dash(e, power, dir, sp, ax, ay);
  } else if (cmd == S2D_CMD_TURN) {
    turn(e, dir, sp);
  } else if (cmd == S2D_CMD_KICK) {
    if (kick(e, power, dir, sp, bax, bay)) e.flags |= S2D_FLAG_KICKED;
  }
  move_object(e.px, e.py, e.vx, e.vy, ax, ay, sp.player_accel_max, sp.player_speed_max, sp.player_decay);
  move_object(e.bx, e.by, e.bvx, e.bvy, bax, bay, sp.ball_accel_max, sp.ball_speed_max, sp.ball_decay);
  collide_ball_player(e, sp);
  update_stamina(e, sp);
  e.cycle += 1u;
}

}  // namespace s2d
