// s2d_fullgame.cuh - the FULLGAME scenario (BASELINE configs[3]): up to 11 v 11, one WARP per match.
//
// Lane l < np owns player l (l < pps: left team, uniform number l+1; else right team); the ball and the referee
// state are replicated in every lane (computed redundantly, so they stay bit-identical across the warp).  Per-team
// and per-match reductions are warp primitives: xor-butterfly shuffles for sums (kick accelerations on the ball,
// collision proposals), ballots for "who kicked / who touched the ball", shuffles to broadcast a partner's position
// in the collision loop.  The player block never leaves registers during the K fused cycles.
//
// Physics per player = the one-player functions of s2d_one_player.cuh (dash / turn / kick / Body_GoToPoint lowering,
// MPObject::_inc, stamina) plus rcssserver's Stadium::collisions for n players; referee subset: goals, ball out ->
// kick-in / corner kick / goal kick, kick-off after a goal, time over.  NOT in the reference (which only ever runs
// one player with the referee off, soccer_2d_env.py:363-369): the spec is include/soccer2d.h.
#pragma once
#include "s2d_scenarios.cuh"

namespace s2d {

constexpr int kFgObsDim = 120;        // 4 ball + 22 x 5 player + 6 referee values (rows are 16-byte multiples)
constexpr int kFgDropBallTime = 100;  // cycles a dead ball waits for the awarded side before play resumes
constexpr int kFgMaxPlayers = 22;
constexpr float kFgFreeKickDist = 9.15f;  // the distance opponents keep from a dead ball
constexpr float kFgOffsideArea = 2.5f;    // offside_active_area_size: a marked player this close to the ball takes part

// HBM layout of N matches with np players each (plane-major; every plane starts 16-byte aligned):
//   PA float4 [N][np] {x, y, vx, vy}            PB float4 [N][np] {body, stamina, effort, recovery}
//   PC float  [N][np] stamina_capacity (plane padded to 16 B)
//   EB float4 [N] ball {x, y, vx, vy}           EF float4 [N] {episode return, player separation bound, offside marks (bits), -}
//   EI uint4  [N] {step_number, cycle, episode, mode | side<<8 | last_touch<<10 | timer<<12 | ball_collided<<20 | done<<21}
//   EJ uint4  [N] {score_l, score_r, collided mask (bit = player), kicked mask}
struct FgLayout {
  int64_t n;
  int np;
  __host__ __device__ size_t pa() const { return 0; }
  __host__ __device__ size_t pb() const { return static_cast<size_t>(n) * np * 16; }
  __host__ __device__ size_t pc() const { return static_cast<size_t>(n) * np * 32; }
  __host__ __device__ size_t eb() const { return pc() + ((static_cast<size_t>(n) * np * 4 + 15) & ~static_cast<size_t>(15)); }
  __host__ __device__ size_t ef() const { return eb() + static_cast<size_t>(n) * 16; }
  __host__ __device__ size_t ei() const { return ef() + static_cast<size_t>(n) * 16; }
  __host__ __device__ size_t ej() const { return ei() + static_cast<size_t>(n) * 16; }
  __host__ __device__ size_t bytes() const { return ej() + static_cast<size_t>(n) * 16; }
};

#ifndef S2D_HOST_EMU

struct Match {
  int step_number;
  uint32_t cycle, episode;
  int mode, side, last_touch, timer;
  int score_l, score_r;
  float ep_return;
  float sep;  // lower bound on the smallest player-player distance (see fg_collisions); 0 = unknown
  uint32_t offside;  // bit = player marked offside at the last pass of its team (all bits from one team)
  bool done_flag;
};

// 4-4-2 kick-off formation of the left team (own half); the right team is the mirror image
__device__ __constant__ float kFgFormX[11] = {-50, -36, -36, -36, -36, -20, -20, -20, -20, -9, -9};
__device__ __constant__ float kFgFormY[11] = {0, -20, -7, 7, 20, -24, -8, 8, 24, -10, 10};

__device__ __forceinline__ float butterfly_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__device__ __forceinline__ float butterfly_max(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}

// kick-off placement of one player (lane): formation spot plus a +-2 m jitter drawn per (episode, kick-off, player)
__device__ __forceinline__ void fg_place_player(Episode& p, const KernelParams& P, uint64_t gid, uint32_t episode,
                                                int lane, int pps, int kick_offs) {
  const bool left = lane < pps;
  const int k = left ? lane : lane - pps;
  const uint4 w = philox4x32_10(P.seed, gid, episode, RNG_RESET,
                                static_cast<uint32_t>(2 + lane) + 32u * (static_cast<uint32_t>(kick_offs) & 0xFFFFu));
  const float jx = u32_to_unit(w.x) * 4.0f - 2.0f, jy = u32_to_unit(w.y) * 4.0f - 2.0f;
  const float fx = kFgFormX[k], fy = kFgFormY[k];
  p.px = (left ? fx : -fx) + jx;
  p.py = (left ? fy : -fy) + jy;
  p.vx = 0.0f;
  p.vy = 0.0f;
  p.body = left ? 0.0f : 180.0f;
}

// Stadium::collisions for np players and the ball, one lane per player.  In a round every object collects the
// positions proposed for it and moves to their average; afterwards whatever collided gets vel *= -0.1 once.
// Returns this lane's bits: 1 = player collided, 2 = player touched the ball; ball_collided is warp-uniform.
//
// `sep` (kept in the match state) is a LOWER BOUND on the smallest distance between two players.  Every cycle it
// shrinks by twice the largest distance a player moved in that cycle (`moved2` = this lane's squared step, the
// referee's placements included); only when it drops below the collision distance is the exact minimum recomputed
// (the O(n^2 / 32) pair loop) - in open play every few cycles instead of every cycle.  The ball is tested against every player every cycle.  If neither test finds an overlap, the ordered
// relaxation rounds - which would change nothing - are skipped, so the results do not depend on this shortcut.
template <class SP>
__device__ __forceinline__ int fg_collisions(Episode& p, bool active, int lane, int np, bool ball_fixed, const SP& sp,
                                             bool& ball_collided, float& sep, float moved2) {
  const unsigned full = 0xffffffffu;
  bool collided = false, ballhit = false, ball_any = false;
  const float r = sp.player_size() + sp.ball_size();
  const float r2 = sp.player_size() + sp.player_size();
  const float h = r2 / 2.0f + kCollideEps;
  ball_collided = false;

  bool ball_overlap = false;
  if (active && !ball_fixed) {
    const float dx = p.bx - p.px, dy = p.by - p.py;
    ball_overlap = dx * dx + dy * dy < r * r;
  }
  {
    // the farthest any player moved this cycle (its own move, or the referee placing it 9.15 m from a dead ball):
    // sqrt of the largest squared step over the lanes, rounded up
    const float v2max = __uint_as_float(__reduce_max_sync(full, __float_as_uint(active ? moved2 : 0.0f)));
    float moved;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(moved) : "f"(v2max));
    sep -= 2.002f * moved + 1.0e-6f;
  }
  bool pairs_close = false;
  if (sep < r2) {  // uniform: the bound has run out, measure the true minimum (lane i looks at (i + d) mod np, d <= np/2)
    float m2 = 3.0e38f;
    int j = lane;
#pragma unroll 1
    for (int d = 1; d <= (np >> 1); ++d) {
      j = j + 1 >= np ? j + 1 - np : j + 1;
      const float xj = __shfl_sync(full, p.px, j), yj = __shfl_sync(full, p.py, j);
      const float ex = p.px - xj, ey = p.py - yj;
      m2 = fminf(m2, ex * ex + ey * ey);
    }
    m2 = active ? m2 : 3.0e38f;
    m2 = __uint_as_float(__reduce_min_sync(full, __float_as_uint(m2)));  // non-negative floats order like their bits
    pairs_close = m2 < r2 * r2;
    sep = sqrtf(m2) * 0.999f;
  }
  if (!pairs_close && !__any_sync(full, ball_overlap)) return 0;

#pragma unroll 1
  for (int round = 0; round < 10; ++round) {
    bool col = false;
    int cnt = 0;
    float sx = 0.0f, sy = 0.0f, bpx = 0.0f, bpy = 0.0f;
    bool bc = false;
#pragma unroll 1
    for (int j = 0; j < np; ++j) {
      const float xj = __shfl_sync(full, p.px, j), yj = __shfl_sync(full, p.py, j);
      if (!active) continue;
      if (j == lane) {
        if (ball_fixed) continue;
        const float dx = p.bx - p.px, dy = p.by - p.py;
        if (dx * dx + dy * dy < r * r) {
          col = collided = ballhit = bc = true;
          const float2 b = ball_back_trace(p.px, p.py, p.bx, p.by, p.bvx, p.bvy, r + kCollideEps);
          bpx = b.x;
          bpy = b.y;
          sx += p.px;
          sy += p.py;
          cnt += 1;
        }
      } else {
        const float ex = p.px - xj, ey = p.py - yj;
        if (ex * ex + ey * ey < r2 * r2) {
          col = collided = true;
          const float mx = (p.px + xj) / 2.0f, my = (p.py + yj) / 2.0f;
          const float d = hypot2(ex, ey);
          float ux, uy;
          if (d < 1.0e-10f) {
            ux = lane < j ? 1.0f : -1.0f;
            uy = 0.0f;
          } else {
            ux = ex / d;
            uy = ey / d;
          }
          sx += mx + ux * h;
          sy += my + uy * h;
          cnt += 1;
        }
      }
    }
    const int bcnt = __popc(__ballot_sync(full, bc));
    const float bsx = butterfly_sum(bpx), bsy = butterfly_sum(bpy);
    if (bcnt) {
      p.bx = bsx / static_cast<float>(bcnt);
      p.by = bsy / static_cast<float>(bcnt);
      ball_any = true;
    }
    if (cnt) {
      p.px = sx / static_cast<float>(cnt);
      p.py = sy / static_cast<float>(cnt);
    }
    if (!__any_sync(full, col)) break;
  }
  sep = 0.0f;  // players were pushed around: measure again next cycle
  if (ball_any) {
    p.bvx *= -0.1f;
    p.bvy *= -0.1f;
  }
  if (collided) {
    p.vx *= -0.1f;
    p.vy *= -0.1f;
  }
  ball_collided = ball_any;
  return (collided ? 1 : 0) | (ballhit ? 2 : 0);
}

// 120-float observation of a match: ball, 22 x {x, y, vx, vy, body}, referee state.  Staged in shared memory and
// written by lanes 0..29 as one float4 each (480 contiguous bytes).
__device__ __forceinline__ void fg_write_obs(float* __restrict__ dst, int64_t env, const Episode& p, const Match& m,
                                             bool active, int lane, int np, int half_time,
                                             float* stage /* [120] */) {
  if (np < kFgMaxPlayers) {  // fewer than 11 a side: the rows of the absent players are zero
    for (int k = lane; k < kFgObsDim; k += 32) stage[k] = 0.0f;
    __syncwarp();
  }
  if (active) {
    float* o = stage + 4 + 5 * lane;
    o[0] = p.px * static_cast<float>(1.0 / 52.5);
    o[1] = p.py * static_cast<float>(1.0 / 34.0);
    o[2] = p.vx;
    o[3] = p.vy;
    o[4] = p.body * static_cast<float>(1.0 / 180.0);
  }
  if (lane == 31) {
    stage[0] = p.bx * static_cast<float>(1.0 / 52.5);
    stage[1] = p.by * static_cast<float>(1.0 / 34.0);
    stage[2] = p.bvx * static_cast<float>(1.0 / 3.0);
    stage[3] = p.bvy * static_cast<float>(1.0 / 3.0);
    stage[114] = static_cast<float>(m.mode);
    stage[115] = static_cast<float>(m.side);
    stage[116] = static_cast<float>(m.score_l);
    stage[117] = static_cast<float>(m.score_r);
    stage[118] = static_cast<float>(m.step_number) / static_cast<float>(2 * half_time);
    stage[119] = 0.0f;
  }
  __syncwarp();
  if (lane < kFgObsDim / 4)
    st_stream(reinterpret_cast<float4*>(dst + env * kFgObsDim) + lane, reinterpret_cast<const float4*>(stage)[lane]);
  __syncwarp();
}

// new match: scores 0, kick-off formation, everybody recovered, kick-off for the left team
template <class SP>
__device__ __forceinline__ void fg_reset(Episode& p, Match& m, const KernelParams& P, const SP& sp, uint64_t gid, int lane,
                                         int pps) {
  m.score_l = 0;
  m.score_r = 0;
  if (lane < 2 * pps) fg_place_player(p, P, gid, m.episode, lane, pps, 0);  // (idle lanes would index past the formation)
  recover(p, sp);
  p.bx = p.by = p.bvx = p.bvy = 0.0f;
  m.episode += 1u;
  m.step_number = 0;
  m.ep_return = 0.0f;
  m.sep = 0.0f;
  m.offside = 0u;
  m.done_flag = false;
  m.mode = S2D_PM_KICK_OFF;
  m.side = S2D_SIDE_LEFT;
  m.timer = 0;
  m.last_touch = S2D_SIDE_UNKNOWN;
}

// The 22 commands of a cycle.  In a match every lane may carry a different command, so a switch over the command
// would make the warp walk dash, turn, kick and go-to-point one after the other, each with its own sincos / atan2 /
// sqrt.  Here the commands are decomposed into the pieces they share, each evaluated ONCE per warp under a vote and
// with per-lane operands:
//   geometry to a reference point  (go-to-point: the target, kick: the ball)  -> distance, relative angle
//   speed                          (turn and go-to-point's turn: inertia)
//   sincos(body + direction)       (dash, go-to-point's dash (direction 0), kick)
// Per lane the arithmetic is the sequence of decode_command + dash_apply / turn / kick (s2d_one_player.cuh), so the
// results are the same bit for bit.  Returns whether this lane kicked (kax, kay = its push on the ball).
template <class SP>
__device__ __forceinline__ bool fg_commands(Episode& p, float4 a, float goto_dist_thr, const SP& sp, const NoiseCtx& nz,
                                            int lane, bool active, bool left, bool may_kick, float& ax, float& ay,
                                            float& kax, float& kay) {
  const unsigned full = 0xffffffffu;
  // the proxy's other body actions become a plain turn or kick first (rare: one vote when nobody uses them)
  if (__any_sync(full, active && a.x >= static_cast<float>(S2D_CMD_TURN_TO_POINT))) {
    if (active && a.x >= static_cast<float>(S2D_CMD_TURN_TO_POINT)) lower_body_action(p, a, sp);
  }
  const int c = active ? static_cast<int>(a.x) : S2D_CMD_NONE;
  const bool is_goto = c == S2D_CMD_GOTO;
  const bool is_kick = c == S2D_CMD_KICK && may_kick;
  const bool user_dash = c == S2D_CMD_DASH, user_turn = c == S2D_CMD_TURN;

  // geometry: player -> reference point
  float dist = 0.0f, rel = 0.0f;
  if (__any_sync(full, is_goto || is_kick)) {
    // (lanes that are not concerned get harmless operands: a zero numerator or denominator, or sqrt(0), would send
    // the whole warp through the slow paths of the IEEE division and square root)
    const float dx = is_goto ? a.y - p.px : is_kick ? p.bx - p.px : 1.0f;
    const float dy = is_goto ? a.z - p.py : is_kick ? p.by - p.py : 0.5f;
    dist = hypot2(dx, dy);
    rel = norm_deg_360(atan2_deg(dy, dx) - p.body);
  }

  // go-to-point decides: nothing (arrived) | turn towards the target | dash straight at it
  bool goto_turn = false, goto_dash = false;
  if (__any_sync(full, is_goto)) {
    const bool on_the_way = is_goto && !(dist < goto_dist_thr);
    const float ratio = goto_dist_thr / (is_goto ? dist : 1.0f);
    float athr = 15.0f;  // asin(ratio) exceeds 15 degrees only for ratio > sin(15 deg) = 0.2588
    if (__any_sync(full, on_the_way && ratio > 0.25f)) {
      const float wide = fmax_(15.0f, atan2_deg(ratio, sqrtf(fmax_(0.0f, 1.0f - ratio * ratio))));
      athr = ratio > 0.25f ? wide : 15.0f;
    }
    goto_turn = on_the_way && fabsf(rel) > athr;
    goto_dash = on_the_way && !goto_turn;
  }

  // turn
  const bool do_turn = user_turn || goto_turn;
  if (__any_sync(full, do_turn)) {
    const float speed = hypot2(do_turn ? p.vx : 1.0f, do_turn ? p.vy : 0.0f);
    const float inertia = 1.0f + sp.inertia_moment() * speed;
    float moment = goto_turn ? clampf(sp.min_moment(), rel * inertia, sp.max_moment()) : user_turn ? a.y : 1.0f;
    moment = clampf(sp.min_moment(), moment, sp.max_moment());
    if (SP::kNoise) moment = moment * (1.0f + sp.player_rand() * u11(noise_block(nz, static_cast<uint32_t>(lane)).z));
    const float body = norm_deg(p.body + moment / inertia);
    p.body = do_turn ? body : p.body;
  }

  // dash direction (go-to-point dashes straight: direction 0) and the one sincos
  const bool do_dash = user_dash || goto_dash;
  float dir = 0.0f, rate = 1.0f;
  if (__any_sync(full, do_dash)) dash_direction(user_dash ? a.z : 0.0f, sp, dir, rate);
  const float user_power = clampf(sp.min_dash_power(), a.y, sp.max_dash_power());
  const bool back = user_dash && user_power < 0.0f;
  float sn = 0.0f, cs = 1.0f;
  if (__any_sync(full, do_dash || is_kick)) {
    const float kick_dir = clampf(sp.min_moment(), a.z, sp.max_moment());
    const float d = is_kick ? kick_dir : back ? dir + 180.0f : dir;
    sincos_deg(p.body + d, sn, cs);
  }

  // dash
  if (__any_sync(full, do_dash)) {
    float power = user_power;
    if (__any_sync(full, goto_dash)) {
      const float v_along = p.vx * cs + p.vy * sn;
      const float need = (dist - v_along) / (goto_dash ? p.effort * sp.dash_power_rate() : 1.0f);
      const float reach = clampf(sp.min_dash_power(), clampf(0.0f, need, a.w), sp.max_dash_power());
      power = goto_dash ? reach : power;
    }
    float need = back ? power * -2.0f : power;
    need = fmin_(need, p.stamina + sp.extra_stamina());
    const float stamina = fmax_(0.0f, p.stamina - need);
    power = back ? need * -0.5f : need;
    float eff = fabsf(p.effort * power * rate * sp.dash_power_rate());
    const float slow = left ? sp.slowness_on_top_for_left_team() : sp.slowness_on_top_for_right_team();
    if (slow != 1.0f && p.py < 0.0f) eff = cold_div(eff, slow);
    p.stamina = do_dash ? stamina : p.stamina;
    ax = do_dash ? eff * cs : 0.0f;
    ay = do_dash ? eff * sn : 0.0f;
  }

  // kick
  bool kicked = false;
  if (__any_sync(full, is_kick)) {
    kicked = is_kick && !(dist > sp.kickable_area());
    const float power = clampf(0.0f, a.y, sp.max_power());
    const float dir_diff = fabsf(rel);
    const float dist_ball = dist - sp.player_size() - sp.ball_size();
    const float eff = power * sp.kick_power_rate() *
                      (1.0f - 0.25f * dir_diff * static_cast<float>(1.0 / 180.0) - 0.25f * dist_ball / sp.kickable_margin());
    float bax = 0.0f, bay = 0.0f;
    bax += eff * cs;
    bay += eff * sn;
    if (SP::kNoise) {
      if (__any_sync(full, kicked)) {
        const uint4 w = noise_block(nz, static_cast<uint32_t>(lane)), w2 = noise_block(nz, 32u + static_cast<uint32_t>(lane));
        const float pos_rate = 0.5f + 0.25f * (dir_diff * static_cast<float>(1.0 / 180.0) + dist_ball / sp.kickable_margin());
        const float speed_rate = 0.5f + 0.5f * (hypot2(p.bvx, p.bvy) / (sp.ball_speed_max() * sp.ball_decay()));
        const float max_rand = sp.kick_rand() * (power / sp.max_power()) * (pos_rate + speed_rate);
        const float mag = u32_to_unit(w.w) * max_rand;
        float ns, nc;
        sincos_deg(u11(w2.x) * 180.0f, ns, nc);
        bax += mag * nc;
        bay += mag * ns;
      }
    }
    kax = kicked ? bax : 0.0f;
    kay = kicked ? bay : 0.0f;
  }
  return kicked;
}

// One cycle of the match.  `a` = this lane's command {cmd, a, b, c}.  Returns done; reward / result are uniform.
template <class SP>
__device__ __forceinline__ bool fg_cycle(Episode& p, Match& m, const KernelParams& P, const SP& sp, uint64_t gid, int lane,
                                         bool active, int np, int half_time, float4 a, float& reward, int& result,
                                         uint32_t& collided_mask, uint32_t& kicked_mask, bool& ball_collided) {
  const unsigned full = 0xffffffffu;
  const int pps = np >> 1;
  const bool left = lane < pps;
  const int my_side = left ? S2D_SIDE_LEFT : S2D_SIDE_RIGHT;
  bool dead = m.mode != S2D_PM_PLAY_ON;
  m.step_number += 1;

  // ---- commands ----
  float ax = 0.0f, ay = 0.0f, kax = 0.0f, kay = 0.0f;
  const NoiseCtx nz{P.seed, gid, m.cycle};
  const bool kicked = fg_commands(p, a, P.goto_dist_thr, sp, nz, lane, active, left, !dead || my_side == m.side, ax, ay, kax, kay);
  const unsigned kick_ballot = __ballot_sync(full, kicked);
  float bax = 0.0f, bay = 0.0f;
  if (kick_ballot) {  // uniform; without a kicker both sums are exactly zero
    bax = butterfly_sum(kax);
    bay = butterfly_sum(kay);
  }
  const unsigned left_lanes = (1u << pps) - 1u;
  const bool kick_l = (kick_ballot & left_lanes) != 0, kick_r = (kick_ballot & ~left_lanes) != 0;
  if (kick_l != kick_r) m.last_touch = kick_l ? S2D_SIDE_LEFT : S2D_SIDE_RIGHT;
  const int mode_at_kick = m.mode;
  if (dead && ((m.side == S2D_SIDE_LEFT && kick_l) || (m.side == S2D_SIDE_RIGHT && kick_r))) {
    m.mode = S2D_PM_PLAY_ON;
    dead = false;
  }
  kicked_mask = kick_ballot;
  // ---- offside marks (OffsideRef): taken at the moment of a pass; uniform branch, only in cycles with a kick ----
  if (kick_ballot && !dead) {
    m.offside = 0u;  // whoever kicks: the old marks are void
    const bool exempt = mode_at_kick == S2D_PM_KICK_IN || mode_at_kick == S2D_PM_CORNER_KICK || mode_at_kick == S2D_PM_GOAL_KICK;
    if (kick_l != kick_r && !exempt) {
      // in the attackers' direction (x mirrored for the right team): beyond the ball, the half-way line and the
      // second-last defender = offside
      const bool att_left = kick_l;
      const float sgn = att_left ? 1.0f : -1.0f;
      const bool defender = active && left != att_left;
      const float v = defender ? sgn * p.px : -3.0e38f;
      const float last = butterfly_max(v);
      const int who = __ffs(__ballot_sync(full, defender && v == last)) - 1;
      const float second = butterfly_max(lane == who ? -3.0e38f : v);
      const float line_x = fmaxf(fmaxf(second, sgn * p.bx), 0.0f);
      const bool marked = active && left == att_left && ((kick_ballot >> lane) & 1u) == 0u && sgn * p.px > line_x;
      m.offside = __ballot_sync(full, marked);
    }
  }

  // ---- move ----
  const float pbx = p.bx, pby = p.by;
  const float ppx = p.px, ppy = p.py;
  if (active)
    move_object<SP::kNoise>(p.px, p.py, p.vx, p.vy, ax, ay, sp.player_accel_max(), sp.player_accel_max2(),
                            sp.player_speed_max(), sp.player_speed_max2(), sp.player_decay(), sp.player_rand(), &nz,
                            static_cast<uint32_t>(lane));
  if (!dead) {
    move_object<SP::kNoise>(p.bx, p.by, p.bvx, p.bvy, bax, bay, sp.ball_accel_max(), sp.ball_accel_max2(),
                            sp.ball_speed_max(), sp.ball_speed_max2(), sp.ball_decay(), sp.ball_rand(), &nz, kBallAgent);
  } else {
    p.bvx = 0.0f;
    p.bvy = 0.0f;
  }

  // ---- dead ball: the side that does not take the kick keeps 9.15 m away (uniform branch; rare) ----
  if (dead && m.mode != S2D_PM_TIME_OVER) {
    const float cx = p.px - p.bx, cy = p.py - p.by;
    const float c2 = cx * cx + cy * cy;
    const bool inside = active && my_side != m.side && c2 < kFgFreeKickDist * kFgFreeKickDist;
    if (__any_sync(full, inside)) {
      if (inside) {
        const float c = sqrtf(c2);
        float ux = left ? -1.0f : 1.0f, uy = 0.0f;
        if (c >= 1.0e-6f) {
          ux = cx / c;
          uy = cy / c;
        }
        p.px = p.bx + ux * kFgFreeKickDist;
        p.py = p.by + uy * kFgFreeKickDist;
        p.vx = 0.0f;
        p.vy = 0.0f;
      }
    }
  }

  // ---- collisions ----
  const float stepx = p.px - ppx, stepy = p.py - ppy;
  const int hit = fg_collisions(p, active, lane, np, dead, sp, ball_collided, m.sep, stepx * stepx + stepy * stepy);
  collided_mask = __ballot_sync(full, (hit & 1) != 0);
  {
    const unsigned touch = __ballot_sync(full, (hit & 2) != 0);
    const bool hit_l = (touch & left_lanes) != 0, hit_r = (touch & ~left_lanes) != 0;
    if (hit_l != hit_r) m.last_touch = hit_l ? S2D_SIDE_LEFT : S2D_SIDE_RIGHT;
    if (m.offside && (touch & ((m.offside & left_lanes) ? ~left_lanes : left_lanes))) m.offside = 0u;  // the defenders got the ball
  }

  // ---- referee (uniform across the warp) ----
  int goal_l = 0, goal_r = 0;
  const float bx_phys = p.bx;
  bool kick_off = false;
  const float line = sp.pitch_half_length() + sp.ball_size();
  const float side_line = sp.pitch_half_width() + sp.ball_size();
  bool offside_called = false;
  if (!dead && m.offside) {  // a marked player within 2.5 m of the ball takes part in play: free kick where it stands
    const float ox = p.px - p.bx, oy = p.py - p.by;
    const unsigned part = __ballot_sync(full, ((m.offside >> lane) & 1u) != 0u && ox * ox + oy * oy < kFgOffsideArea * kFgOffsideArea);
    if (part) {
      const int who = __ffs(part) - 1;
      const float fx = __shfl_sync(full, p.px, who), fy = __shfl_sync(full, p.py, who);
      m.mode = S2D_PM_FREE_KICK;
      m.side = who < pps ? S2D_SIDE_RIGHT : S2D_SIDE_LEFT;
      m.timer = 0;
      p.bx = clampf(-sp.pitch_half_length(), fx, sp.pitch_half_length());
      p.by = clampf(-sp.pitch_half_width(), fy, sp.pitch_half_width());
      p.bvx = 0.0f;
      p.bvy = 0.0f;
      offside_called = true;
    }
  }
  if (offside_called) {
    // (the ball was re-placed inside the pitch: nothing else to rule on this cycle)
  } else if (!dead && !(fabsf(p.bx) > line || fabsf(p.by) > side_line)) {
    // ball inside the field: every ruling below needs it beyond a line, so there is nothing to decide (the common case)
  } else if (!dead) {
    const float bx = p.bx, by = p.by;
    const float post = sp.goal_width() * 0.5f + sp.goal_post_radius();
    if (bx > line && !(pbx > line)) {
      const float yc = pby + (by - pby) * ((line - pbx) / (bx - pbx));
      goal_l = fabsf(yc) <= post;
    } else if (bx < -line && !(pbx < -line)) {
      const float yc = pby + (by - pby) * ((-line - pbx) / (bx - pbx));
      goal_r = fabsf(yc) <= post;
    }
    if (goal_l || goal_r) {
      if (goal_l) m.score_l += 1;
      else m.score_r += 1;
      kick_off = true;
      m.mode = S2D_PM_KICK_OFF;
      m.side = goal_l ? S2D_SIDE_RIGHT : S2D_SIDE_LEFT;  // the conceding side kicks off
      m.timer = 0;
      m.last_touch = S2D_SIDE_UNKNOWN;
    } else if (fabsf(bx) > line) {  // over a goal line outside the goal: corner kick or goal kick
      const int defending = bx > 0.0f ? S2D_SIDE_RIGHT : S2D_SIDE_LEFT;
      const float sx = bx > 0.0f ? 1.0f : -1.0f, sy = by > 0.0f ? 1.0f : -1.0f;
      if (m.last_touch == defending) {
        m.mode = S2D_PM_CORNER_KICK;
        m.side = defending == S2D_SIDE_LEFT ? S2D_SIDE_RIGHT : S2D_SIDE_LEFT;
        p.bx = sx * (sp.pitch_half_length() - 1.0f);
        p.by = sy * (sp.pitch_half_width() - 1.0f);
      } else {
        m.mode = S2D_PM_GOAL_KICK;
        m.side = defending;
        p.bx = sx * (sp.pitch_half_length() - 5.5f);
        p.by = sy * 9.16f;
      }
      p.bvx = 0.0f;
      p.bvy = 0.0f;
      m.timer = 0;
    } else if (fabsf(by) > side_line) {
      m.mode = S2D_PM_KICK_IN;
      m.side = m.last_touch == S2D_SIDE_LEFT ? S2D_SIDE_RIGHT : S2D_SIDE_LEFT;
      p.bx = clampf(-sp.pitch_half_length(), bx, sp.pitch_half_length());
      p.by = by > 0.0f ? sp.pitch_half_width() : -sp.pitch_half_width();
      p.bvx = 0.0f;
      p.bvy = 0.0f;
      m.timer = 0;
    }
  } else {
    m.timer += 1;
    if (m.timer >= kFgDropBallTime) {
      m.mode = S2D_PM_PLAY_ON;
      m.timer = 0;
    }
  }
  if (kick_off) {
    m.sep = 0.0f;
    if (active) fg_place_player(p, P, gid, m.episode, lane, pps, m.score_l + m.score_r);
    p.bx = p.by = p.bvx = p.bvy = 0.0f;
  }
  if (m.mode != S2D_PM_PLAY_ON) m.offside = 0u;  // marks live only while play goes on
  if (active) update_stamina(p, sp);
  m.cycle += 1u;
  reward = static_cast<float>(goal_l - goal_r) * 10.0f + (bx_phys - pbx) * 0.01f;
  const bool done = m.step_number >= 2 * half_time;
  result = !done ? S2D_RESULT_NONE : m.score_l > m.score_r ? 1 : m.score_r > m.score_l ? 2 : 3;
  if (done) m.mode = S2D_PM_TIME_OVER;
  return done;
}

__device__ __forceinline__ void fg_load(const KernelParams& P, const FgLayout& L, int64_t env, int lane, bool active,
                                        Episode& p, Match& m) {
  const char* base = static_cast<const char*>(P.state);
  p = Episode{};
  if (active) {
    const int64_t idx = env * L.np + lane;
    const float4 a = ld_stream(reinterpret_cast<const float4*>(base + L.pa()) + idx);
    const float4 b = ld_stream(reinterpret_cast<const float4*>(base + L.pb()) + idx);
    p.px = a.x; p.py = a.y; p.vx = a.z; p.vy = a.w;
    p.body = b.x; p.stamina = b.y; p.effort = b.z; p.recovery = b.w;
    p.capacity = *(reinterpret_cast<const float*>(base + L.pc()) + idx);
  }
  const float4 ball = *(reinterpret_cast<const float4*>(base + L.eb()) + env);
  const float4 ef = *(reinterpret_cast<const float4*>(base + L.ef()) + env);
  const uint4 ei = *(reinterpret_cast<const uint4*>(base + L.ei()) + env);
  const uint4 ej = *(reinterpret_cast<const uint4*>(base + L.ej()) + env);
  p.bx = ball.x; p.by = ball.y; p.bvx = ball.z; p.bvy = ball.w;
  m.ep_return = ef.x;
  m.sep = ef.y;
  m.offside = __float_as_uint(ef.z);
  m.step_number = static_cast<int>(ei.x); m.cycle = ei.y; m.episode = ei.z;
  m.mode = ei.w & 0xff; m.side = (ei.w >> 8) & 3; m.last_touch = (ei.w >> 10) & 3; m.timer = (ei.w >> 12) & 0xff;
  m.done_flag = (ei.w >> 21) & 1;
  m.score_l = static_cast<int>(ej.x); m.score_r = static_cast<int>(ej.y);
}

__device__ __forceinline__ void fg_store(const KernelParams& P, const FgLayout& L, int64_t env, int lane, bool active,
                                         const Episode& p, const Match& m, uint32_t collided_mask, uint32_t kicked_mask,
                                         bool ball_collided) {
  char* base = static_cast<char*>(P.state);
  if (active) {
    const int64_t idx = env * L.np + lane;
    st_stream(reinterpret_cast<float4*>(base + L.pa()) + idx, make_float4(p.px, p.py, p.vx, p.vy));
    st_stream(reinterpret_cast<float4*>(base + L.pb()) + idx, make_float4(p.body, p.stamina, p.effort, p.recovery));
    *(reinterpret_cast<float*>(base + L.pc()) + idx) = p.capacity;
  }
  if (lane == 0) {
    *(reinterpret_cast<float4*>(base + L.eb()) + env) = make_float4(p.bx, p.by, p.bvx, p.bvy);
    *(reinterpret_cast<float4*>(base + L.ef()) + env) = make_float4(m.ep_return, m.sep, __uint_as_float(m.offside), 0.0f);
    const uint32_t packed = static_cast<uint32_t>(m.mode) | (static_cast<uint32_t>(m.side) << 8) |
                            (static_cast<uint32_t>(m.last_touch) << 10) | (static_cast<uint32_t>(m.timer) << 12) |
                            (ball_collided ? 1u << 20 : 0u) | (m.done_flag ? 1u << 21 : 0u);
    *(reinterpret_cast<uint4*>(base + L.ei()) + env) =
        make_uint4(static_cast<uint32_t>(m.step_number), m.cycle, m.episode, packed);
    *(reinterpret_cast<uint4*>(base + L.ej()) + env) =
        make_uint4(static_cast<uint32_t>(m.score_l), static_cast<uint32_t>(m.score_r), collided_mask, kicked_mask);
  }
}

// heterogeneous players: the block copies the type table to shared memory and every lane points at its player's row
template <class SP>
__device__ __forceinline__ void fg_bind_player_type(SP& sp, const KernelParams& P, int lane) {
  if constexpr (SP::kHetero) {
    __shared__ float s_types[S2D_MAX_PLAYER_TYPES * PT_ROW];
    for (int k = threadIdx.x; k < S2D_MAX_PLAYER_TYPES * PT_ROW; k += blockDim.x) s_types[k] = __ldg(P.player_types + k);
    __syncthreads();
    sp.row = s_types + PT_ROW * P.type_of[lane];
  }
}

#ifndef S2D_FG_MIN_BLOCKS
#define S2D_FG_MIN_BLOCKS 8
#endif
constexpr int kFgBlock = 128;  // 4 matches per block

// K lockstep cycles of every match; actions float4 [N][K][np].  NP = 22 is the 11 v 11 instantiation (player count,
// team masks and row strides become immediates; otherwise the compiler keeps re-reading them from the constant bank
// under register pressure); NP = 0 takes the player count at run time.
template <int VAR, int NP>
__global__ void __launch_bounds__(kFgBlock, S2D_FG_MIN_BLOCKS) fullgame_step_kernel(const __grid_constant__ KernelParams P, const int K,
                                                                 const int np_runtime, const int half_time) {
  const int np = NP ? NP : np_runtime;
  using SP = typename VariantSP<VAR>::type;
  SP sp(P.cc);
  __shared__ __align__(16) float s_stage[kFgBlock / 32][kFgObsDim];
  const int lane = threadIdx.x & 31;
  fg_bind_player_type(sp, P, lane);
  const int64_t env = static_cast<int64_t>(blockIdx.x) * (kFgBlock / 32) + (threadIdx.x >> 5);
  if (env >= P.num_envs) return;  // whole warp leaves together
  const FgLayout L{P.num_envs, np};
  const bool active = lane < np;
  const uint64_t gid = static_cast<uint64_t>(P.env_id_offset + env);
  float* stage = s_stage[threadIdx.x >> 5];

  Episode p;
  Match m;
  fg_load(P, L, env, lane, active, p, m);
  uint32_t collided_mask = 0, kicked_mask = 0;
  bool ball_collided = false;
  float reward_sum = 0.0f;
  uint32_t any_done = 0, last_result = 0;
  // the lane's commands: K rows np apart.  The next row is fetched while this cycle computes (its latency would
  // otherwise sit in front of every cycle: nothing can start before the command is known).
  const float4* act = static_cast<const float4*>(P.actions) + (env * K) * np + (active ? lane : 0);
  float4 a_next = __ldg(act);
#pragma unroll 1
  for (int k = K; k > 0; --k) {
    const float4 a = active ? a_next : make_float4(0.f, 0.f, 0.f, 0.f);
    act += np;
    if (k > 1) a_next = __ldg(act);
    float rw;
    int rs;
    const bool done = fg_cycle(p, m, P, sp, gid, lane, active, np, half_time, a, rw, rs, collided_mask, kicked_mask,
                               ball_collided);
    reward_sum += rw;
    m.ep_return += rw;
    if (done) {
      any_done = 1;
      last_result = static_cast<uint32_t>(rs);
      if (lane == 0 && !m.done_flag) {  // (a finished match stepped on with auto_reset off is tallied once)
        unsigned long long* slot = P.stats + static_cast<size_t>(env % kStatSlots) * kStatWords;
        atomicAdd(slot + ST_EPISODES, 1ull);
        atomicAdd(slot + (rs == 1 ? ST_GOALS : rs == 2 ? ST_OUTS : ST_TIMEOUTS), 1ull);
        atomicAdd(slot + ST_EP_STEPS, static_cast<unsigned long long>(m.step_number));
        atomicAdd(reinterpret_cast<double*>(slot + ST_RETURN), static_cast<double>(m.ep_return));
      }
      if (P.terminal_obs) fg_write_obs(P.terminal_obs, env, p, m, active, lane, np, half_time, stage);
      if (P.auto_reset) {
        fg_reset(p, m, P, sp, gid, lane, np >> 1);
        collided_mask = kicked_mask = 0;
        ball_collided = false;
      } else {
        m.done_flag = true;
      }
    }
  }
  fg_store(P, L, env, lane, active, p, m, collided_mask, kicked_mask, ball_collided);
  fg_write_obs(P.obs, env, p, m, active, lane, np, half_time, stage);
  if (lane == 0) {
    P.reward[env] = reward_sum;
    P.done[env] = static_cast<uint8_t>(any_done);
    P.result[env] = static_cast<uint8_t>(last_result);
  }
}

template <bool HETERO>
__global__ void __launch_bounds__(kFgBlock) fullgame_reset_kernel(const __grid_constant__ KernelParams P,
                                                                  const uint8_t* __restrict__ mask, const int np,
                                                                  const int half_time) {
  __shared__ __align__(16) float s_stage[kFgBlock / 32][kFgObsDim];
  const int lane = threadIdx.x & 31;
  using SP = typename std::conditional<HETERO, HeteroSP<false>, RuntimeSP>::type;
  SP sp(P.cc);
  fg_bind_player_type(sp, P, lane);
  const int64_t env = static_cast<int64_t>(blockIdx.x) * (kFgBlock / 32) + (threadIdx.x >> 5);
  if (env >= P.num_envs) return;
  if (mask && !mask[env]) return;
  const FgLayout L{P.num_envs, np};
  const bool active = lane < np;
  Episode p;
  Match m;
  fg_load(P, L, env, lane, active, p, m);
  fg_reset(p, m, P, sp, static_cast<uint64_t>(P.env_id_offset + env), lane, np >> 1);
  fg_store(P, L, env, lane, active, p, m, 0u, 0u, false);
  fg_write_obs(P.obs, env, p, m, active, lane, np, half_time, s_stage[threadIdx.x >> 5]);
  if (lane == 0) {
    P.reward[env] = 0.0f;
    P.done[env] = 0;
    P.result[env] = 0;
  }
}

#endif  // !S2D_HOST_EMU

}  // namespace s2d
