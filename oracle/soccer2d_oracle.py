"""CPU ORACLE (pure Python, float64) for the Soccer2DEnv.step/reset hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under gym-soccer-2d-env_b200/ may import this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use oracle/.

What it restates and how each part is pinned
--------------------------------------------
(1) ENV CONTRACT - transcribed line by line from the reference (paths under /root/reference):
      action decode        sample_environments/reach_ball_env.py:53-85
      observation build    sample_environments/reach_ball_env.py:87-111
      reward / done / info sample_environments/reach_ball_env.py:113-161
      reset distribution   sample_environments/reach_ball_env.py:170-218
      reset/step sequence  soccer_2d_env.py:179-269, reach_ball_env.py:163-168
    PINNED: tests/golden/reach_ball_contract.json was produced by executing the reference's own
    reach_ball_env.py (tests/golden/make_golden.py); tests/test_oracle_contract.py checks this file
    against every vector in it.  The geometry helper (pyrusgeom 0.1.2, requirements.txt:6) is not
    installed offline and is restated below (`norm_deg`, `atan2_deg`, ...).

(2) PHYSICS - rcssserver (github.com/CLSFramework/rcssserver, release "latest", un-pinned:
    scripts/download-rcssserver.sh:30) and the proxy's action lowering
    (github.com/clsframework/soccer-simulation-proxy, "latest": scripts/download-proxy.sh:30) are external
    binaries that are not in /root/reference and not available offline.  The functions below restate
    rcssserver's published model (src/player.cpp Player::dash/turn/kick/updateStamina,
    src/object.cpp MPObject::_inc, src/stadium.cpp Stadium::collisions / movePlayer / recover,
    src/referee.cpp BallOut/Goal subset) from its manual and source AS REMEMBERED.
    **PARITY UNPINNED for this part**: there is no rcssserver binary, test or golden vector to check it
    against; every decision that could not be verified is marked "DECISION" below and listed in DESIGN.md.

Arithmetic: Python floats (IEEE double) like upstream rcssserver; the values that cross the proto
boundary (service.proto declares every scalar `float`) are rounded to float32 with `f32()` exactly where
the reference's env would see them (obs, reward inputs, reset placement).  The C restatement
(oracle/s2d_oracle.c) follows the same operation order; its float32 build is the bit-exact mirror of the
CUDA kernels.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

# ----------------------------------------------------------------------------------------------
# Parameters (names = proto ServerParam / PlayerType fields, idl/service.proto:1435-1732;
# values = rcssserver defaults, SURVEY.md Appendix A.1 - they are NOT in the reference tree)
# ----------------------------------------------------------------------------------------------


@dataclass
class ServerParam:
    pitch_half_length: float = 52.5
    pitch_half_width: float = 34.0
    goal_width: float = 14.02
    goal_post_radius: float = 0.06
    ball_size: float = 0.085
    ball_decay: float = 0.94
    ball_rand: float = 0.05
    ball_speed_max: float = 3.0
    ball_accel_max: float = 2.7
    player_size: float = 0.3
    player_decay: float = 0.4
    player_rand: float = 0.1
    player_speed_max: float = 1.05
    player_accel_max: float = 1.0
    dash_power_rate: float = 0.006
    inertia_moment: float = 5.0
    min_dash_power: float = 0.0
    max_dash_power: float = 100.0
    min_dash_angle: float = -180.0
    max_dash_angle: float = 180.0
    dash_angle_step: float = 1.0
    side_dash_rate: float = 0.4
    back_dash_rate: float = 0.7
    min_power: float = -100.0
    max_power: float = 100.0
    min_moment: float = -180.0
    max_moment: float = 180.0
    kick_power_rate: float = 0.027
    kickable_margin: float = 0.7
    kick_rand: float = 0.1
    stamina_max: float = 8000.0
    stamina_inc_max: float = 45.0
    extra_stamina: float = 50.0
    stamina_capacity: float = 130600.0
    recover_init: float = 1.0
    recover_min: float = 0.5
    recover_dec: float = 0.002
    recover_dec_thr: float = 0.3
    effort_init: float = 1.0
    effort_max: float = 1.0
    effort_min: float = 0.6
    effort_dec: float = 0.005
    effort_dec_thr: float = 0.3
    effort_inc: float = 0.01
    effort_inc_thr: float = 0.6
    slowness_on_top_for_left_team: float = 1.0
    slowness_on_top_for_right_team: float = 1.0
    noise: bool = False  # "noise off" = player_rand = ball_rand = kick_rand = 0 (north_star)

    def as_f32(self) -> "ServerParam":
        """The same constants rounded to float32 - what a proto ServerParam message (all `float` fields,
        idl/service.proto:1435-1662) and the C ABI's S2DServerParam carry."""
        out = ServerParam()
        for k, v in self.__dict__.items():
            setattr(out, k, f32(v) if isinstance(v, float) else v)
        return out


# play-mode codes = proto GameModeType (idl/service.proto:267-301); sides = proto Side (:88-92)
PM_BeforeKickOff, PM_TimeOver, PM_PlayOn, PM_KickOff, PM_KickIn, PM_FreeKick, PM_CornerKick, PM_GoalKick, PM_AfterGoal = range(9)
SIDE_UNKNOWN, SIDE_LEFT, SIDE_RIGHT = 0, 1, 2

# command codes = PlayerAction oneof tags we lower (idl/service.proto:1291-1308: dash=1, turn=2, kick=3,
# body_go_to_point=16); 0 = no body command this cycle (e.g. Body_HoldBall without a kickable ball)
CMD_NONE, CMD_DASH, CMD_TURN, CMD_KICK, CMD_GOTO = 0, 1, 2, 3, 4

RESULT_NONE, RESULT_GOAL, RESULT_OUT, RESULT_TIMEOUT = 0, 1, 2, 3
RESULT_NAMES = (None, "Goal", "Out", "Timeout")  # reach_ball_env.py:126,140,145,150


def f32(x: float) -> float:
    """Round to IEEE float32 (what a proto `float` field stores)."""
    return struct.unpack("f", struct.pack("f", x))[0]


# ----------------------------------------------------------------------------------------------
# Geometry (pyrusgeom Vector2D / AngleDeg restated; call sites reach_ball_env.py:89-96,119-124)
# ----------------------------------------------------------------------------------------------


def norm_deg(d: float) -> float:
    """AngleDeg normalisation: into [-180, 180] (both ends kept)."""
    if d < -360.0 or 360.0 < d:
        d = math.fmod(d, 360.0)
    if d < -180.0:
        d += 360.0
    if d > 180.0:
        d -= 360.0
    return d


def atan2_deg(y: float, x: float) -> float:
    """Vector2D.th(): 0 for the zero vector (DECISION: exact-zero test as in librcsc; pyrusgeom may use an
    epsilon - immaterial away from the origin)."""
    if x == 0.0 and y == 0.0:
        return 0.0
    return math.degrees(math.atan2(y, x))


def polar(r: float, deg: float):
    rad = math.radians(deg)
    return r * math.cos(rad), r * math.sin(rad)


def hypot2(x: float, y: float) -> float:
    return math.sqrt(x * x + y * y)


def clamp(lo: float, x: float, hi: float) -> float:
    return max(lo, min(x, hi))


# ----------------------------------------------------------------------------------------------
# Counter-based RNG: Philox4x32-10 (Salmon et al., SC'11).  key = 64-bit seed; counter =
# (env_id_lo, env_id_hi, index, (purpose << 24) | sub).  Identical in s2d_oracle.c and the CUDA kernels.
# ----------------------------------------------------------------------------------------------

PHILOX_M0, PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
PHILOX_W0, PHILOX_W1 = 0x9E3779B9, 0xBB67AE85
M32 = 0xFFFFFFFF

RNG_RESET, RNG_BALLVEL, RNG_ACTION, RNG_NOISE = 0, 1, 2, 3


def philox4x32(counter, key):
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + PHILOX_W0) & M32
        k1 = (k1 + PHILOX_W1) & M32
    return c0, c1, c2, c3


def rng_block(seed: int, env_id: int, index: int, purpose: int, sub: int = 0):
    return philox4x32((env_id & M32, (env_id >> 32) & M32, index & M32, ((purpose << 24) | sub) & M32),
                      (seed & M32, (seed >> 32) & M32))


def u32_to_int(u: int, lo: int, hi: int) -> int:
    """Uniform integer in [lo, hi] inclusive (multiply-shift; random.randint's role, reach_ball_env.py:173-178)."""
    return lo + ((u * (hi - lo + 1)) >> 32)


def u32_to_unit(u: int) -> float:
    """Uniform in [0, 1) with 24 bits (random.random's role, reach_ball_env.py:205)."""
    return (u >> 8) * (1.0 / 16777216.0)


# ----------------------------------------------------------------------------------------------
# Physics objects
# ----------------------------------------------------------------------------------------------


@dataclass
class Ball:
    x: float = 0.0
    y: float = 0.0
    vx: float = 0.0
    vy: float = 0.0
    ax: float = 0.0
    ay: float = 0.0
    collided: bool = False


@dataclass
class Player:
    x: float = 0.0
    y: float = 0.0
    vx: float = 0.0
    vy: float = 0.0
    ax: float = 0.0
    ay: float = 0.0
    body: float = 0.0  # degrees, [-180, 180]
    stamina: float = 8000.0
    effort: float = 1.0
    recovery: float = 1.0
    capacity: float = 130600.0
    side: int = SIDE_LEFT
    collided: bool = False
    kicked: bool = False


def player_recover(p: Player, sp: ServerParam) -> None:
    """Trainer `(recover)` = DoRecover (idl/service.proto:1407): Stadium::recoveryPlayers -> Player::recoverAll."""
    p.stamina = sp.stamina_max
    p.recovery = sp.recover_init
    p.effort = sp.effort_max
    p.capacity = sp.stamina_capacity


def cmd_dash(p: Player, power: float, direction: float, sp: ServerParam) -> None:
    """Player::dash (rcssserver src/player.cpp)."""
    power = clamp(sp.min_dash_power, power, sp.max_dash_power)
    direction = clamp(sp.min_dash_angle, direction, sp.max_dash_angle)
    if sp.dash_angle_step > 1.0e-10:
        # rint(): round-half-to-even, so 22.5 -> 22 and 67.5 -> 68
        direction = sp.dash_angle_step * round(direction / sp.dash_angle_step)
    back = power < 0.0
    need = power * -2.0 if back else power
    need = min(need, p.stamina + sp.extra_stamina)
    p.stamina = max(0.0, p.stamina - need)
    power = need / -2.0 if back else need
    ad = math.fabs(direction)
    if ad > 90.0:
        dir_rate = sp.back_dash_rate - ((sp.back_dash_rate - sp.side_dash_rate) * (1.0 - (ad - 90.0) / 90.0))
    else:
        dir_rate = sp.side_dash_rate + ((1.0 - sp.side_dash_rate) * (1.0 - ad / 90.0))
    dir_rate = clamp(0.0, dir_rate, 1.0)
    eff = math.fabs(p.effort * power * dir_rate * sp.dash_power_rate)
    if p.y < 0.0:
        eff /= sp.slowness_on_top_for_left_team if p.side == SIDE_LEFT else sp.slowness_on_top_for_right_team
    if back:
        direction += 180.0
    ax, ay = polar(eff, p.body + direction)
    p.ax += ax
    p.ay += ay


def cmd_turn(p: Player, moment: float, sp: ServerParam) -> None:
    """Player::turn: actual turn shrinks with speed (inertia_moment)."""
    moment = clamp(sp.min_moment, moment, sp.max_moment)
    speed = hypot2(p.vx, p.vy)
    p.body = norm_deg(p.body + moment / (1.0 + sp.inertia_moment * speed))


def kickable(p: Player, b: Ball, sp: ServerParam) -> bool:
    return hypot2(b.x - p.x, b.y - p.y) <= sp.player_size + sp.ball_size + sp.kickable_margin


def cmd_kick(p: Player, b: Ball, power: float, direction: float, sp: ServerParam, play_mode: int = PM_PlayOn) -> bool:
    """Player::kick + Stadium::kickTaken (accelerations of several kickers add up).
    DECISION: power is clamped to [0, max_power] (old servers allowed min_power < 0)."""
    power = clamp(0.0, power, sp.max_power)
    direction = clamp(sp.min_moment, direction, sp.max_moment)
    if play_mode in (PM_BeforeKickOff, PM_AfterGoal, PM_TimeOver):
        return False
    dx, dy = b.x - p.x, b.y - p.y
    dist = hypot2(dx, dy)
    if dist > sp.player_size + sp.ball_size + sp.kickable_margin:
        return False
    dir_diff = math.fabs(norm_deg(atan2_deg(dy, dx) - p.body))  # degrees
    dist_ball = dist - sp.player_size - sp.ball_size
    eff = power * sp.kick_power_rate * (1.0 - 0.25 * dir_diff / 180.0 - 0.25 * dist_ball / sp.kickable_margin)
    ax, ay = polar(eff, p.body + direction)
    b.ax += ax
    b.ay += ay
    p.kicked = True
    return True


def lower_goto(p: Player, tx: float, ty: float, dist_thr: float, max_power: float, sp: ServerParam):
    """Body_GoToPoint (idl/service.proto:684-688) lowered by the proxy (librcsc Body_GoToPoint) to one
    turn or dash.  DECISION (simplified librcsc rule): no omni-dash, no stamina saving;
    dir_thr = 15 degrees; dash power = what reaches the target this cycle without overshoot."""
    dx, dy = tx - p.x, ty - p.y
    dist = hypot2(dx, dy)
    if dist < dist_thr:
        return CMD_NONE, 0.0, 0.0
    ang = norm_deg(atan2_deg(dy, dx) - p.body)
    ratio = dist_thr / dist
    thr = max(15.0, math.degrees(math.atan2(ratio, math.sqrt(max(0.0, 1.0 - ratio * ratio)))))  # asin
    if math.fabs(ang) > thr:
        speed = hypot2(p.vx, p.vy)
        return CMD_TURN, 0.0, clamp(sp.min_moment, ang * (1.0 + sp.inertia_moment * speed), sp.max_moment)
    # first-cycle travel of a forward dash = |vel| along body + effort*power*rate
    rad = math.radians(p.body)
    v_along = p.vx * math.cos(rad) + p.vy * math.sin(rad)
    need = (dist - v_along) / (p.effort * sp.dash_power_rate)
    return CMD_DASH, clamp(0.0, need, max_power), 0.0


CMD_TURN_TO_POINT, CMD_TURN_TO_BALL, CMD_TURN_TO_ANGLE, CMD_KICK_ONE_STEP, CMD_STOP_BALL, CMD_INTERCEPT = 5, 6, 7, 8, 9, 10
CMD_TACKLE, CMD_CATCH, CMD_SMART_KICK = 11, 12, 13  # tackle / catch: FULLGAME only (tests/fullgame_twin.py)


def inertia_factor(decay: float, n: int) -> float:
    """(1 - d^n) / (1 - d): how far an object drifts in n cycles per unit of velocity"""
    return (1.0 - decay ** n) / (1.0 - decay)


def lower_body_action(p: Player, b: Ball, c: int, a1: float, a2: float, a3: float, sp: ServerParam):
    """Body_TurnToPoint / Body_TurnToBall / Body_TurnToAngle / Body_KickOneStep (force mode) / Body_StopBall /
    Body_Intercept (idl/service.proto:742-780; the spec is in include/soccer2d.h) lowered, as the proxy does with
    librcsc's rules, to (cmd, a1, a2, a3) of the basic vocabulary.  Values handed on are proto floats (f32)."""
    if c in (CMD_TURN_TO_POINT, CMD_TURN_TO_BALL, CMD_TURN_TO_ANGLE):
        if c == CMD_TURN_TO_ANGLE:
            ang = norm_deg(norm_deg(a1) - p.body)
        else:
            n = int(clamp(0.0, float(round_half_even(a1 if c == CMD_TURN_TO_BALL else a3)), 63.0))
            tx, ty = a1, a2
            if c == CMD_TURN_TO_BALL:
                fb = inertia_factor(sp.ball_decay, n)
                tx, ty = b.x + b.vx * fb, b.y + b.vy * fb
            fp = inertia_factor(sp.player_decay, n)
            mx, my = p.x + p.vx * fp, p.y + p.vy * fp
            ang = norm_deg(atan2_deg(ty - my, tx - mx) - p.body)
        speed = hypot2(p.vx, p.vy)
        return CMD_TURN, f32(clamp(sp.min_moment, ang * (1.0 + sp.inertia_moment * speed), sp.max_moment)), 0.0, 0.0
    if c in (CMD_KICK_ONE_STEP, CMD_STOP_BALL):
        dx, dy = b.x - p.x, b.y - p.y
        dist = hypot2(dx, dy)
        if dist > sp.player_size + sp.ball_size + sp.kickable_margin:
            return CMD_NONE, 0.0, 0.0, 0.0
        wx = wy = 0.0
        if c == CMD_KICK_ONE_STEP:
            first_speed = clamp(0.0, a3, sp.ball_speed_max)
            th = math.radians(atan2_deg(a2 - b.y, a1 - b.x))
            wx, wy = first_speed * math.cos(th), first_speed * math.sin(th)
        ax, ay = wx - b.vx, wy - b.vy
        acc = hypot2(ax, ay)
        dir_diff = math.fabs(norm_deg(atan2_deg(dy, dx) - p.body))
        dist_ball = dist - sp.player_size - sp.ball_size
        rate = sp.kick_power_rate * (1.0 - 0.25 * dir_diff / 180.0 - 0.25 * dist_ball / sp.kickable_margin)
        return CMD_KICK, f32(clamp(0.0, acc / rate, sp.max_power)), f32(norm_deg(atan2_deg(ay, ax) - p.body)), 0.0
    if c == CMD_SMART_KICK:  # Body_SmartKick (idl/service.proto:690-695): release if one kick can do it, else stage
        dx, dy = b.x - p.x, b.y - p.y
        dist = hypot2(dx, dy)
        if dist > sp.player_size + sp.ball_size + sp.kickable_margin:
            return CMD_NONE, 0.0, 0.0, 0.0
        first_speed = clamp(0.0, a3, sp.ball_speed_max)
        th = math.radians(atan2_deg(a2 - b.y, a1 - b.x))
        ax, ay = first_speed * math.cos(th) - b.vx, first_speed * math.sin(th) - b.vy
        acc = hypot2(ax, ay)
        dir_diff = math.fabs(norm_deg(atan2_deg(dy, dx) - p.body))
        dist_ball = dist - sp.player_size - sp.ball_size
        rate = sp.kick_power_rate * (1.0 - 0.25 * dir_diff / 180.0 - 0.25 * dist_ball / sp.kickable_margin)
        if not acc <= sp.max_power * rate:
            nx, ny = p.x + p.vx, p.y + p.vy
            th = math.radians(atan2_deg(a2 - ny, a1 - nx))
            d = sp.player_size + sp.ball_size + 0.3 * sp.kickable_margin
            sax, say = ((nx + d * math.cos(th)) - b.x) - b.vx, ((ny + d * math.sin(th)) - b.y) - b.vy
            if hypot2(sax, say) >= 0.05:  # (else it is staged already: release with what max_power gives)
                ax, ay = sax, say
                acc = hypot2(ax, ay)
        return CMD_KICK, f32(clamp(0.0, acc / rate, sp.max_power)), f32(norm_deg(atan2_deg(ay, ax) - p.body)), 0.0
    if c == CMD_INTERCEPT:
        reach0 = 0.8 * (sp.player_size + sp.ball_size + sp.kickable_margin)
        fb = fp = 0.0
        pwb = pwp = 1.0
        tx, ty = b.x, b.y
        for t in range(1, 31):
            fb += pwb
            fp += pwp
            pwb *= sp.ball_decay
            pwp *= sp.player_decay
            tx, ty = b.x + b.vx * fb, b.y + b.vy * fb
            mx, my = p.x + p.vx * fp, p.y + p.vy * fp
            reach = reach0 + (t - 1) * sp.player_speed_max
            if (tx - mx) ** 2 + (ty - my) ** 2 <= reach * reach:
                break
        return CMD_GOTO, f32(tx), f32(ty), 100.0
    return c, a1, a2, a3


def round_half_even(x: float) -> float:
    return float(round(x))  # Python rounds halves to even, like rint


def obj_inc(o, accel_max: float, speed_max: float, decay: float) -> None:
    """MPObject::_inc (rcssserver src/object.cpp), noise off, no wind."""
    if o.ax != 0.0 or o.ay != 0.0:
        a = hypot2(o.ax, o.ay)
        if a > accel_max:
            s = accel_max / a
            o.ax *= s
            o.ay *= s
        o.vx += o.ax
        o.vy += o.ay
        v = hypot2(o.vx, o.vy)
        if v > speed_max:
            s = speed_max / v
            o.vx *= s
            o.vy *= s
    o.x += o.vx
    o.y += o.vy
    o.vx *= decay
    o.vy *= decay
    o.ax = 0.0
    o.ay = 0.0


COLLIDE_EPS = 1.0e-6  # DECISION: upstream EPS is 1e-10 in double; 1e-6 is representable next to 0.385 in f32


COLLISION_MIDPOINT, COLLISION_BACKTRACE = 0, 1  # include/soccer2d.h "Collision models"


def collisions(ball: Ball, players: list, sp: ServerParam, model: int = COLLISION_MIDPOINT) -> None:
    """Stadium::collisions: <=10 relaxation rounds; every object moves to the AVERAGE of the
    positions proposed for it in a round; afterwards each object that collided gets vel *= -0.1 once.
    DECISIONS: ball-player: the ball is moved back along its own velocity to the touching distance and the
    player keeps its place; if the ball is (numerically) not moving, or the back-trace has no solution, the
    symmetric player-player rule is used.  Coincident centres separate along +x (noise off, so no random
    direction).
    model = COLLISION_BACKTRACE (the other reading of the pair rule; SURVEY A.5 marks it uncertain): EVERY colliding object
    backs up along its own velocity until it touches the other one - both players of a pair, and the player of a
    ball-player contact as well as the ball."""
    ball.collided = False
    for p in players:
        p.collided = False
    n = len(players)
    for _ in range(10):
        col = False
        bsx = bsy = 0.0
        bcnt = 0
        acc = [[0.0, 0.0, 0] for _ in range(n)]
        for i in range(n):
            pi = players[i]
            r = sp.player_size + sp.ball_size
            dx, dy = ball.x - pi.x, ball.y - pi.y
            if dx * dx + dy * dy < r * r:
                col = True
                ball.collided = True
                pi.collided = True
                nx, ny = _ball_back_trace(ball, pi, r + COLLIDE_EPS)
                bsx += nx
                bsy += ny
                bcnt += 1
                qx, qy = (pi.x, pi.y) if model == COLLISION_MIDPOINT else \
                    _trace_back(pi.x, pi.y, pi.vx, pi.vy, ball.x, ball.y, r + COLLIDE_EPS, -1.0)
                acc[i][0] += qx
                acc[i][1] += qy
                acc[i][2] += 1
            for j in range(i + 1, n):
                pj = players[j]
                r2 = sp.player_size + sp.player_size
                dx, dy = pi.x - pj.x, pi.y - pj.y
                if dx * dx + dy * dy < r2 * r2:
                    col = True
                    pi.collided = True
                    pj.collided = True
                    if model == COLLISION_BACKTRACE:
                        for a, o, side in ((i, j, 1.0), (j, i, -1.0)):
                            pa, po = players[a], players[o]
                            qx, qy = _trace_back(pa.x, pa.y, pa.vx, pa.vy, po.x, po.y, r2 + COLLIDE_EPS, side)
                            acc[a][0] += qx
                            acc[a][1] += qy
                            acc[a][2] += 1
                        continue
                    mx, my = (pi.x + pj.x) / 2.0, (pi.y + pj.y) / 2.0
                    d = hypot2(dx, dy)
                    if d < 1.0e-10:
                        ux, uy = 1.0, 0.0
                    else:
                        ux, uy = dx / d, dy / d
                    h = r2 / 2.0 + COLLIDE_EPS
                    acc[i][0] += mx + ux * h
                    acc[i][1] += my + uy * h
                    acc[i][2] += 1
                    acc[j][0] += mx - ux * h
                    acc[j][1] += my - uy * h
                    acc[j][2] += 1
        if bcnt:
            ball.x, ball.y = bsx / bcnt, bsy / bcnt
        for i in range(n):
            if acc[i][2]:
                players[i].x, players[i].y = acc[i][0] / acc[i][2], acc[i][1] / acc[i][2]
        if not col:
            break
    if ball.collided:
        ball.vx *= -0.1
        ball.vy *= -0.1
    for p in players:
        if p.collided:
            p.vx *= -0.1
            p.vy *= -0.1


def _trace_back(x, y, vx, vy, fx, fy, r: float, side: float = 1.0):
    """Point on the incoming line of the object at (x, y) with velocity (vx, vy) - pos - t*vel_dir, t >= 0 - at distance
    r from the object at (fx, fy); at rest or when the line misses: straight out along the line of centres; coincident
    centres separate along x (side)."""
    # the velocity the object arrived with (vel was already decayed by _inc; direction is what matters)
    v = hypot2(vx, vy)
    dx, dy = x - fx, y - fy
    if v > 1.0e-10:
        ux, uy = vx / v, vy / v
        # |d - t u|^2 = r^2  ->  t^2 - 2 t (d.u) + |d|^2 - r^2 = 0, larger root moves the object back out
        du = dx * ux + dy * uy
        disc = du * du - (dx * dx + dy * dy - r * r)
        if disc >= 0.0:
            t = du + math.sqrt(disc)
            if t >= 0.0:
                return x - t * ux, y - t * uy
    d = hypot2(dx, dy)
    if d < 1.0e-10:
        return fx + side * r, fy
    return fx + dx / d * r, fy + dy / d * r


def _ball_back_trace(ball: Ball, p: Player, r: float):
    return _trace_back(ball.x, ball.y, ball.vx, ball.vy, p.x, p.y, r)


def update_stamina(p: Player, sp: ServerParam) -> None:
    """Player::updateStamina incl. stamina_capacity bookkeeping."""
    if p.stamina <= sp.recover_dec_thr * sp.stamina_max:
        if p.recovery > sp.recover_min:
            p.recovery -= sp.recover_dec
        if p.recovery < sp.recover_min:
            p.recovery = sp.recover_min
    if p.stamina <= sp.effort_dec_thr * sp.stamina_max:
        if p.effort > sp.effort_min:
            p.effort -= sp.effort_dec
        if p.effort < sp.effort_min:
            p.effort = sp.effort_min
    if p.stamina >= sp.effort_inc_thr * sp.stamina_max:
        if p.effort < sp.effort_max:
            p.effort += sp.effort_inc
            if p.effort > sp.effort_max:
                p.effort = sp.effort_max
    inc = min(p.recovery * sp.stamina_inc_max, sp.stamina_max - p.stamina)
    if sp.stamina_capacity >= 0.0:
        if inc > p.capacity:
            inc = p.capacity
    p.stamina += inc
    if sp.stamina_capacity >= 0.0:
        p.capacity = max(0.0, p.capacity - inc)


# ----------------------------------------------------------------------------------------------
# ReachBall: the reference's only scenario (configs[0], configs[1])
# ----------------------------------------------------------------------------------------------

ACT_DISCRETE, ACT_CONTINUOUS, ACT_TURNING, ACT_COMMAND = 0, 1, 2, 3
BALLVEL_MAX_TRIES = 64  # bound on the rejection loop of reach_ball_env.py:204-212 (fallback: speed 0)


@dataclass
class ReachBallConfig:
    change_ball_position: bool = True
    change_ball_velocity: bool = False
    ball_position_x: float = 0.0
    ball_position_y: float = 0.0
    ball_speed: float = 0.0
    ball_direction: float = 0.0
    min_distance_to_ball: float = 5.0
    max_steps: int = 200
    use_continuous_action: bool = True
    action_space_size: int = 16
    use_turning: bool = False
    seed: int = 0
    sp: ServerParam = field(default_factory=ServerParam)
    collision_model: int = 0  # COLLISION_MIDPOINT / COLLISION_BACKTRACE

    @property
    def action_mode(self) -> int:
        if self.use_continuous_action:
            return ACT_TURNING if self.use_turning else ACT_CONTINUOUS
        return ACT_DISCRETE


def decode_action(cfg: ReachBallConfig, action, u: float):
    """reach_ball_env.py:53-85 -> (cmd, power, relative_direction as the proto float32 holds it).
    `u` replaces np.random.rand() at :71."""
    if cfg.use_continuous_action:
        if cfg.use_turning:
            a = [clamp(-1.0, float(v), 1.0) for v in action]
            turn_prob, turn_angle, dash_prob, dash_angle = a
            e0, e1 = math.exp(dash_prob), math.exp(turn_prob)
            p0 = e0 / (e0 + e1)
            if u < p0:  # :71-72: "turn_selected" is tested against the DASH logit's softmax weight
                return CMD_TURN, 0.0, f32(turn_angle * 180.0)
            return CMD_DASH, 100.0, f32(dash_angle * 180.0)
        a = float(action[0]) if hasattr(action, "__len__") else float(action)
        return CMD_DASH, 100.0, f32(a * 180.0)
    a = int(action)
    return CMD_DASH, 100.0, f32((a * 360.0 / cfg.action_space_size) % 360.0 - 180.0)


def build_obs(bx, by, bvx, bvy, px, py, body):
    """reach_ball_env.py:87-111 on float32-quantised proto fields; returns 10 Python floats (float64)."""
    ball_speed = hypot2(bvx, bvy)
    ball_direction = norm_deg(atan2_deg(bvy, bvx))
    pb = norm_deg(body)
    player_to_ball = norm_deg(atan2_deg(by - py, bx - px))
    body_to_ball = norm_deg(player_to_ball - pb)
    return [body_to_ball / 180.0, pb / 180.0, px / 52.5, py / 34.0, bx / 52.5, by / 34.0,
            ball_speed / 3.0, ball_direction / 360.0, bvx / 3.0, bvy / 3.0]


def check_trainer(cfg: ReachBallConfig, mem_dist: float, mem_ang: float, step_number: int, bx, by, px, py, body):
    """reach_ball_env.py:113-161 -> (done, reward, result_code, new_mem_dist, new_mem_ang)."""
    dist = hypot2(bx - px, by - py)
    diff = norm_deg(norm_deg(atan2_deg(by - py, bx - px)) - norm_deg(body))
    done, reward, result = False, 0.0, RESULT_NONE
    reward += mem_dist - dist
    reward += (math.fabs(norm_deg(mem_ang)) - math.fabs(diff)) / 180.0
    if dist < cfg.min_distance_to_ball:
        done = True
        reward += 10.0
        result = RESULT_GOAL
    if math.fabs(px) > 52.5 or math.fabs(py) > 34.0:
        done = True
        reward -= -10.0  # reference quirk (:144): leaving the pitch ADDS 10
        result = RESULT_OUT
    if step_number > cfg.max_steps:
        done = True
        reward -= 5.0
        result = RESULT_TIMEOUT
    return done, reward, result, dist, diff


class ReachBallOracle:
    """One ReachBall episode stream = what Soccer2DEnv.step/reset would return if rcssserver followed the
    restated physics.  `env_id` is the GLOBAL env index (RNG key), so shards reproduce the same episodes."""

    def __init__(self, cfg: ReachBallConfig, env_id: int = 0, auto_reset: bool = True):
        self.cfg = cfg
        self.sp = cfg.sp
        self.env_id = env_id
        self.auto_reset = auto_reset
        self.episode = 0  # index of the NEXT reset's random block
        self.cycle = 0  # server cycle: monotonic across episodes, like rcssserver's
        self.step_number = 0
        self.mem_dist = 0.0
        self.mem_ang = 0.0
        self.ep_return = 0.0
        self.ball = Ball()
        self.player = Player()
        self.travel_factor = (1.0 - 0.96 ** cfg.max_steps) / (1.0 - 0.96)  # :207 (0.96, not ball_decay)

    # -- reach_ball_env.py:170-218 with Philox in place of the Mersenne Twister --------------------
    def sample_reset(self):
        cfg = self.cfg
        w = rng_block(cfg.seed, self.env_id, self.episode, RNG_RESET, 0)
        px = u32_to_int(w[0], -50, 50)
        py = u32_to_int(w[1], -30, 30)
        body = u32_to_int(w[2], 0, 360)
        if cfg.change_ball_position:
            w2 = rng_block(cfg.seed, self.env_id, self.episode, RNG_RESET, 1)
            bx = float(u32_to_int(w[3], -50, 50))
            by = float(u32_to_int(w2[0], -30, 30))
        else:
            bx, by = float(cfg.ball_position_x), float(cfg.ball_position_y)
        if cfg.change_ball_velocity:
            speed, d = 0.0, 0.0
            for t in range(BALLVEL_MAX_TRIES):
                wv = rng_block(cfg.seed, self.env_id, self.episode, RNG_BALLVEL, t >> 1)
                s = u32_to_unit(wv[2 * (t & 1)]) * 3.0
                dd = float(u32_to_int(wv[2 * (t & 1) + 1], 0, 360))
                travel = s * self.travel_factor
                tx, ty = polar(travel, dd)
                if math.fabs(bx + tx) <= 52.5 and math.fabs(by + ty) <= 34.0:
                    speed, d = s, dd
                    break
        else:
            speed, d = float(cfg.ball_speed), float(cfg.ball_direction)
        bvx, bvy = polar(speed, d)
        return float(px), float(py), float(body), bx, by, f32(bvx), f32(bvy)

    def reset(self):
        """soccer_2d_env.py:179-224 + reach_ball_env.py:163-168: place, recover, ONE server cycle with the
        player idle (Body_HoldBall; DECISION: a no-op - the ball is not kickable in practice), observe, prime."""
        px, py, body, bx, by, bvx, bvy = self.sample_reset()
        self.episode += 1
        self.step_number = 0
        self.ep_return = 0.0
        b, p = self.ball, self.player
        b.x, b.y, b.vx, b.vy, b.ax, b.ay = bx, by, bvx, bvy, 0.0, 0.0
        p.x, p.y, p.vx, p.vy, p.ax, p.ay = px, py, 0.0, 0.0, 0.0, 0.0
        p.body = norm_deg(body)
        player_recover(p, self.sp)
        self._simulate_cycle(CMD_NONE, 0.0, 0.0)
        obs = self.observe()
        _, _, _, self.mem_dist, self.mem_ang = self._check()
        return obs

    def _simulate_cycle(self, cmd, power, direction):
        sp, p, b = self.sp, self.player, self.ball
        if cmd == CMD_DASH:
            cmd_dash(p, power, direction, sp)
        elif cmd == CMD_TURN:
            cmd_turn(p, direction, sp)
        obj_inc(p, sp.player_accel_max, sp.player_speed_max, sp.player_decay)
        obj_inc(b, sp.ball_accel_max, sp.ball_speed_max, sp.ball_decay)
        collisions(b, [p], sp, getattr(self.cfg, "collision_model", COLLISION_MIDPOINT))
        update_stamina(p, sp)
        self.cycle += 1

    def quantised(self):
        b, p = self.ball, self.player
        return f32(b.x), f32(b.y), f32(b.vx), f32(b.vy), f32(p.x), f32(p.y), f32(p.body)

    def observe(self):
        return build_obs(*self.quantised())

    def _check(self):
        bx, by, _, _, px, py, body = self.quantised()
        return check_trainer(self.cfg, self.mem_dist, self.mem_ang, self.step_number, bx, by, px, py, body)

    def step(self, action):
        """soccer_2d_env.py:226-269.  Returns (obs, reward, done, result_code, terminal_obs|None); with
        auto_reset the returned obs is the first observation of the next episode (VecEnv convention)."""
        self.step_number += 1
        u = u32_to_unit(rng_block(self.cfg.seed, self.env_id, self.cycle, RNG_ACTION, 0)[0])
        cmd, power, direction = decode_action(self.cfg, action, u)
        self._simulate_cycle(cmd, power, direction)
        obs = self.observe()
        done, reward, result, self.mem_dist, self.mem_ang = self._check()
        self.ep_return += reward
        term = None
        if done and self.auto_reset:
            term = obs
            obs = self.reset()
        return obs, reward, done, result, term


# ----------------------------------------------------------------------------------------------
# Shoot: 1v0 shoot-on-goal (BASELINE configs[2]).  NOT in the reference; the contract is specified in
# include/soccer2d.h.  Second restatement (next to oracle/s2d_oracle.c) so that the two can be checked
# against each other.
# ----------------------------------------------------------------------------------------------


@dataclass
class ShootConfig:
    change_ball_position: bool = True
    change_ball_velocity: bool = False
    ball_position_x: float = 0.0
    ball_position_y: float = 0.0
    ball_speed: float = 0.0
    ball_direction: float = 0.0
    max_steps: int = 200
    action_space_size: int = 24
    kick_actions: int = 8
    goto_dist_thr: float = 0.5
    seed: int = 0
    sp: ServerParam = field(default_factory=ServerParam)
    # ReachBallOracle.sample_reset reads these
    min_distance_to_ball: float = 5.0
    collision_model: int = 0  # COLLISION_MIDPOINT / COLLISION_BACKTRACE


def check_shoot(cfg, mem_pb, mem_bg, step_number, bx, by, px, py, prev_bx, prev_by):
    """-> (done, reward, result, d_player_ball, d_ball_goal); all positions as the proto float32 holds them."""
    sp = cfg.sp
    d_pb = hypot2(bx - px, by - py)
    d_bg = hypot2(sp.pitch_half_length - bx, 0.0 - by)
    reward = (mem_pb - d_pb) * 0.2 + (mem_bg - d_bg)
    line = sp.pitch_half_length + sp.ball_size
    post = sp.goal_width / 2.0 + sp.goal_post_radius
    goal = False
    if bx > line and not prev_bx > line:
        yc = prev_by + (by - prev_by) * ((line - prev_bx) / (bx - prev_bx))
        goal = math.fabs(yc) <= post
    out = (not goal) and (math.fabs(bx) > line or math.fabs(by) > sp.pitch_half_width + sp.ball_size)
    done, result = False, RESULT_NONE
    if goal:
        done, reward, result = True, reward + 10.0, RESULT_GOAL
    elif out:
        done, reward, result = True, reward - 10.0, RESULT_OUT
    elif step_number > cfg.max_steps:
        done, reward, result = True, reward - 5.0, RESULT_TIMEOUT
    return done, reward, result, d_pb, d_bg


class ShootOracle(ReachBallOracle):
    """One Shoot episode stream.  Actions: ('discrete', a) or a 4-tuple command (cmd, a, b, c)."""

    def _dirs(self, a, n):
        return f32((a * 360.0 / n) % 360.0 - 180.0)

    def decode(self, action):
        cfg = self.cfg
        if isinstance(action, (int,)) or (hasattr(action, "dtype") and action.ndim == 0):
            a = int(action)
            n_dash = cfg.action_space_size - cfg.kick_actions
            if a < n_dash:
                return CMD_DASH, 100.0, self._dirs(a, n_dash)
            return CMD_KICK, 100.0, self._dirs(a - n_dash, cfg.kick_actions)
        c, a1, a2, a3 = (float(v) for v in action)
        if c >= CMD_TURN_TO_POINT:
            c, a1, a2, a3 = lower_body_action(self.player, self.ball, int(c), a1, a2, a3, self.sp)
        c = int(c)
        if c in (CMD_DASH, CMD_KICK):
            return c, a1, a2
        if c == CMD_TURN:
            return c, 0.0, a1
        if c == CMD_GOTO:
            return lower_goto(self.player, a1, a2, f32(cfg.goto_dist_thr), a3, self.sp)
        return CMD_NONE, 0.0, 0.0

    def _simulate_cycle(self, cmd, power, direction):
        sp, p, b = self.sp, self.player, self.ball
        p.kicked = False
        if cmd == CMD_DASH:
            cmd_dash(p, power, direction, sp)
        elif cmd == CMD_TURN:
            cmd_turn(p, direction, sp)
        elif cmd == CMD_KICK:
            cmd_kick(p, b, power, direction, sp)
        obj_inc(p, sp.player_accel_max, sp.player_speed_max, sp.player_decay)
        obj_inc(b, sp.ball_accel_max, sp.ball_speed_max, sp.ball_decay)
        collisions(b, [p], sp, getattr(self.cfg, "collision_model", COLLISION_MIDPOINT))
        update_stamina(p, sp)
        self.cycle += 1

    def _check(self):
        bx, by, _, _, px, py, _ = self.quantised()
        return check_shoot(self.cfg, self.mem_dist, self.mem_ang, self.step_number, bx, by, px, py,
                           f32(self._prev_ball[0]), f32(self._prev_ball[1]))

    def reset(self):
        px, py, body, bx, by, bvx, bvy = self.sample_reset()
        self.episode += 1
        self.step_number = 0
        self.ep_return = 0.0
        b, p = self.ball, self.player
        b.x, b.y, b.vx, b.vy, b.ax, b.ay = bx, by, bvx, bvy, 0.0, 0.0
        p.x, p.y, p.vx, p.vy, p.ax, p.ay = px, py, 0.0, 0.0, 0.0, 0.0
        p.body = norm_deg(body)
        player_recover(p, self.sp)
        self._prev_ball = (b.x, b.y)
        self._simulate_cycle(CMD_NONE, 0.0, 0.0)
        obs = self.observe()
        _, _, _, self.mem_dist, self.mem_ang = self._check()
        return obs

    def step(self, action):
        self.step_number += 1
        cmd, power, direction = self.decode(action)
        self._prev_ball = (self.ball.x, self.ball.y)
        self._simulate_cycle(cmd, power, direction)
        obs = self.observe()
        done, reward, result, self.mem_dist, self.mem_ang = self._check()
        self.ep_return += reward
        term = None
        if done and self.auto_reset:
            term = obs
            obs = self.reset()
        return obs, reward, done, result, term
