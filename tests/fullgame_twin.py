"""A second, independent restatement of ONE cycle of the FULLGAME scenario, written from the spec in include/soccer2d.h
with the building blocks of the Python oracle (oracle/soccer2d_oracle.py): commands, moves, dead-ball clearance, n-body
collisions, offside, referee, stamina, reward.  Test infrastructure: tests/test_oracle_c.py steps the C oracle's f64
build and this twin from the same state and compares them one cycle ahead (a match is chaotic, so not a whole
trajectory).  Pure Python, double precision, noise off."""
from __future__ import annotations

import math

from oracle import soccer2d_oracle as O

FORM = [(-50, 0), (-36, -20), (-36, -7), (-36, 7), (-36, 20), (-20, -24), (-20, -8), (-20, 8), (-20, 24), (-9, -10), (-9, 10)]
DROP_BALL_TIME, FREE_KICK_DIST, OFFSIDE_AREA = 100, 9.15, 2.5
AFTER_GOAL_WAIT, TACKLE_CYCLES, CATCH_BAN, RNG_TACKLE = 50, 10, 5, 5
PM_AfterGoal = 8


class Match:
    """state in the layout of s2do_get_state_fg: np x 12 player values, 5 ball values, 12 referee values"""

    def __init__(self, vec, np_players):
        self.n = n = np_players
        self.players = []
        for j in range(n):
            v = vec[12 * j:12 * j + 12]
            p = O.Player(x=v[0], y=v[1], vx=v[2], vy=v[3], body=v[4], stamina=v[5], effort=v[6], recovery=v[7], capacity=v[8],
                         side=int(v[11]))
            p.collided, p.kicked = bool(v[9]), bool(v[10])
            self.players.append(p)
        k = 12 * n
        self.ball = O.Ball(x=vec[k], y=vec[k + 1], vx=vec[k + 2], vy=vec[k + 3])
        self.ball.collided = bool(vec[k + 4])
        (self.step_number, self.cycle, self.episode, self.mode, self.side, self.timer, self.score_l, self.score_r,
         self.last_touch) = (int(x) for x in vec[k + 5:k + 14])
        self.ep_return, self.done_flag, self.offside = float(vec[k + 14]), int(vec[k + 15]), int(vec[k + 16])
        self.tackle = [0] * n       # cycles each player still lies on the ground after a tackle
        self.catch_ban = [0, 0]     # left / right goalkeeper

    def set_extra(self, extra):
        """[np tackle counters, catch ban left, catch ban right] (s2do_get_extra_fg)"""
        self.tackle = [int(x) for x in extra[:self.n]]
        self.catch_ban = [int(extra[self.n]), int(extra[self.n + 1])]

    def extra(self):
        return [float(x) for x in self.tackle + self.catch_ban]

    def vector(self):
        out = []
        for p in self.players:
            out += [p.x, p.y, p.vx, p.vy, p.body, p.stamina, p.effort, p.recovery, p.capacity, float(p.collided), float(p.kicked),
                    float(p.side)]
        b = self.ball
        out += [b.x, b.y, b.vx, b.vy, float(b.collided)]
        out += [self.step_number, self.cycle, self.episode, self.mode, self.side, self.timer, self.score_l, self.score_r,
                self.last_touch, self.ep_return, self.done_flag, self.offside]
        return [float(x) for x in out]


def place_formation(m: Match, seed: int, gid: int):
    pps = m.n // 2
    kick_offs = (m.score_l + m.score_r) & 0xFFFF
    for i, p in enumerate(m.players):
        left = i < pps
        fx, fy = FORM[i if left else i - pps]
        w = O.rng_block(seed, gid, m.episode, O.RNG_RESET, (2 + i) + 32 * kick_offs)
        jx, jy = O.u32_to_unit(w[0]) * 4.0 - 2.0, O.u32_to_unit(w[1]) * 4.0 - 2.0
        p.x, p.y = (fx if left else -fx) + jx, (fy if left else -fy) + jy
        p.vx = p.vy = p.ax = p.ay = 0.0
        p.body = 0.0 if left else 180.0
    b = m.ball
    b.x = b.y = b.vx = b.vy = b.ax = b.ay = 0.0


def decode(p, ball, a, goto_thr, sp):
    c, a1, a2, a3 = float(a[0]), float(a[1]), float(a[2]), float(a[3])
    if c >= O.CMD_TURN_TO_POINT:
        c, a1, a2, a3 = O.lower_body_action(p, ball, int(c), a1, a2, a3, sp)
    c = int(c)
    if c in (O.CMD_DASH, O.CMD_KICK):
        return c, a1, a2
    if c == O.CMD_TURN:
        return c, 0.0, a1
    if c == O.CMD_GOTO:
        return O.lower_goto(p, a1, a2, goto_thr, a3, sp)
    return O.CMD_NONE, 0.0, 0.0


def cycle(m: Match, actions, sps, sp, seed: int, gid: int, goto_thr: float, half_time: int, collision_model: int = 0):
    """one cycle; `actions` [np][4]; `sps[j]` = player j's ServerParam (its player type written over `sp`).
    Returns (reward, done, result)."""
    n, b = m.n, m.ball
    dead = m.mode != O.PM_PlayOn
    m.step_number += 1
    # ---- commands ----
    stopped = m.mode == PM_AfterGoal  # the server's clock stands still: only turns work
    pps = n // 2
    b.ax = b.ay = 0.0
    kick_l = kick_r = False
    caught = -1
    for j, p in enumerate(m.players):
        p.kicked = False
        raw = int(actions[j][0])
        keeper = 0 if j == 0 else 1 if j == pps else -1
        if not stopped and keeper >= 0 and m.catch_ban[keeper] > 0:
            m.catch_ban[keeper] -= 1
        if not stopped and m.tackle[j] > 0:  # on the ground after a tackle
            m.tackle[j] -= 1
            continue
        if raw == O.CMD_TACKLE:
            if stopped:
                continue
            th = math.radians(p.body)
            dx, dy = b.x - p.x, b.y - p.y
            rx, ry = dx * math.cos(th) + dy * math.sin(th), dy * math.cos(th) - dx * math.sin(th)
            m.tackle[j] = TACKLE_CYCLES
            ok = False
            if rx > 0.0:
                fail = (rx / 2.0) ** 6 + (abs(ry) / 1.25) ** 6
                if fail < 1.0:
                    ok = O.u32_to_unit(O.rng_block(seed, gid, m.cycle, RNG_TACKLE, j)[0]) < 1.0 - fail
            if ok and not dead:
                d = O.clamp(-180.0, float(actions[j][1]), 180.0)
                eff = 100.0 * (1.0 - abs(d) / 180.0) * 0.027 * (1.0 - 0.5 * abs(O.atan2_deg(ry, rx)) / 180.0)
                t2 = math.radians(p.body + d)
                b.ax += eff * math.cos(t2)
                b.ay += eff * math.sin(t2)
                p.kicked = True
                if p.side == O.SIDE_LEFT:
                    kick_l = True
                else:
                    kick_r = True
            continue
        if raw == O.CMD_CATCH:
            if stopped or dead or keeper < 0 or m.catch_ban[keeper] > 0:
                continue
            d = O.clamp(-180.0, float(actions[j][1]), 180.0)
            th = math.radians(p.body + d)
            dx, dy = b.x - p.x, b.y - p.y
            rx, ry = dx * math.cos(th) + dy * math.sin(th), dy * math.cos(th) - dx * math.sin(th)
            edge = sp.pitch_half_length - 16.5
            in_area = (b.x <= -edge if keeper == 0 else b.x >= edge) and abs(b.y) <= 20.16
            if 0.0 <= rx <= 1.2 and abs(ry) <= 0.5 and in_area:
                caught = j
                m.catch_ban[keeper] = CATCH_BAN
            continue
        cmd, power, direction = decode(p, b, actions[j], goto_thr, sps[j])
        if stopped and cmd != O.CMD_TURN:
            continue
        if cmd == O.CMD_DASH:
            O.cmd_dash(p, power, direction, sps[j])
        elif cmd == O.CMD_TURN:
            O.cmd_turn(p, direction, sps[j])
        elif cmd == O.CMD_KICK and (not dead or p.side == m.side):
            if O.cmd_kick(p, b, power, direction, sps[j]):
                if p.side == O.SIDE_LEFT:
                    kick_l = True
                else:
                    kick_r = True
    if kick_l != kick_r:
        m.last_touch = O.SIDE_LEFT if kick_l else O.SIDE_RIGHT
    mode_at_kick = m.mode
    if dead and ((m.side == O.SIDE_LEFT and kick_l) or (m.side == O.SIDE_RIGHT and kick_r)):
        m.mode, dead = O.PM_PlayOn, False
    if caught >= 0:  # the goalkeeper holds the ball: free kick for its side
        keeper_side = m.players[caught].side
        m.mode, m.side, m.timer, m.last_touch, dead = O.PM_FreeKick, keeper_side, 0, keeper_side, True
    # ---- offside marks ----
    if (kick_l or kick_r) and not dead:
        m.offside = 0
        exempt = mode_at_kick in (O.PM_KickIn, O.PM_CornerKick, O.PM_GoalKick)
        if kick_l != kick_r and not exempt:
            att = O.SIDE_LEFT if kick_l else O.SIDE_RIGHT
            sgn = 1.0 if kick_l else -1.0
            defenders = sorted((sgn * q.x for q in m.players if q.side != att), reverse=True)
            second = defenders[1] if len(defenders) > 1 else -3.0e38
            line_x = max(second, sgn * b.x, 0.0)
            for j, q in enumerate(m.players):
                if q.side == att and not q.kicked and sgn * q.x > line_x:
                    m.offside |= 1 << j
    # ---- move ----
    pbx, pby = b.x, b.y
    for j, p in enumerate(m.players):
        if stopped:
            p.vx = p.vy = p.ax = p.ay = 0.0
        else:
            O.obj_inc(p, sp.player_accel_max, sp.player_speed_max, sps[j].player_decay)
    if not dead:
        O.obj_inc(b, sp.ball_accel_max, sp.ball_speed_max, sp.ball_decay)
    else:
        b.vx = b.vy = b.ax = b.ay = 0.0
    if caught >= 0:
        b.x, b.y = m.players[caught].x, m.players[caught].y
    # ---- kick-off: everybody in its own half ----
    if m.mode == O.PM_KickOff:
        for p in m.players:
            if (p.x > 0.0) if p.side == O.SIDE_LEFT else (p.x < 0.0):
                p.x = -sp.player_size if p.side == O.SIDE_LEFT else sp.player_size
                p.vx = p.vy = 0.0
    # ---- dead-ball clearance ----
    if dead and m.mode not in (O.PM_TimeOver, PM_AfterGoal):
        for p in m.players:
            if p.side == m.side:
                continue
            cx, cy = p.x - b.x, p.y - b.y
            c2 = cx * cx + cy * cy
            if c2 < FREE_KICK_DIST * FREE_KICK_DIST:
                c = math.sqrt(c2)
                ux, uy = ((-1.0 if p.side == O.SIDE_LEFT else 1.0), 0.0) if c < 1.0e-6 else (cx / c, cy / c)
                p.x, p.y, p.vx, p.vy = b.x + ux * FREE_KICK_DIST, b.y + uy * FREE_KICK_DIST, 0.0, 0.0
    # ---- collisions (a dead ball takes no part) ----
    if stopped:  # nothing moves, so nothing is pushed apart either
        touched = [False] * n
        b.collided = False
        for p in m.players:
            p.collided = False
    else:
        touched = collide(m, sp, dead, collision_model)
    hit_l = any(t and p.side == O.SIDE_LEFT for t, p in zip(touched, m.players))
    hit_r = any(t and p.side == O.SIDE_RIGHT for t, p in zip(touched, m.players))
    if hit_l != hit_r:
        m.last_touch = O.SIDE_LEFT if hit_l else O.SIDE_RIGHT
    if m.offside:
        marked_left = any((m.offside >> j) & 1 and p.side == O.SIDE_LEFT for j, p in enumerate(m.players))
        if (hit_r if marked_left else hit_l):
            m.offside = 0
    # ---- referee ----
    goal_l = goal_r = 0
    bx_phys = b.x
    line, side_line = sp.pitch_half_length + sp.ball_size, sp.pitch_half_width + sp.ball_size
    called = False
    if not dead and m.offside:
        for j, p in enumerate(m.players):
            if (m.offside >> j) & 1 and (p.x - b.x) ** 2 + (p.y - b.y) ** 2 < OFFSIDE_AREA * OFFSIDE_AREA:
                m.mode, m.timer = O.PM_FreeKick, 0
                m.side = O.SIDE_RIGHT if p.side == O.SIDE_LEFT else O.SIDE_LEFT
                b.x = O.clamp(-sp.pitch_half_length, p.x, sp.pitch_half_length)
                b.y = O.clamp(-sp.pitch_half_width, p.y, sp.pitch_half_width)
                b.vx = b.vy = 0.0
                called = True
                break
    if called:
        pass
    elif not dead:
        bx, by = b.x, b.y
        post = sp.goal_width / 2.0 + sp.goal_post_radius
        if bx > line and not pbx > line:
            goal_l = int(abs(pby + (by - pby) * ((line - pbx) / (bx - pbx))) <= post)
        elif bx < -line and not pbx < -line:
            goal_r = int(abs(pby + (by - pby) * ((-line - pbx) / (bx - pbx))) <= post)
        if goal_l or goal_r:
            if goal_l:
                m.score_l += 1
            else:
                m.score_r += 1
            m.mode, m.timer, m.last_touch = PM_AfterGoal, 0, O.SIDE_UNKNOWN
            m.side = O.SIDE_LEFT if goal_l else O.SIDE_RIGHT  # AfterGoal_ + the scoring side
            b.vx = b.vy = 0.0
        elif abs(bx) > line:
            defending = O.SIDE_RIGHT if bx > 0.0 else O.SIDE_LEFT
            sx, sy = (1.0 if bx > 0.0 else -1.0), (1.0 if by > 0.0 else -1.0)
            if m.last_touch == defending:
                m.mode = O.PM_CornerKick
                m.side = O.SIDE_RIGHT if defending == O.SIDE_LEFT else O.SIDE_LEFT
                b.x, b.y = sx * (sp.pitch_half_length - 1.0), sy * (sp.pitch_half_width - 1.0)
            else:
                m.mode, m.side = O.PM_GoalKick, defending
                b.x, b.y = sx * (sp.pitch_half_length - 5.5), sy * 9.16
            b.vx = b.vy = 0.0
            m.timer = 0
        elif abs(by) > side_line:
            m.mode = O.PM_KickIn
            m.side = O.SIDE_RIGHT if m.last_touch == O.SIDE_LEFT else O.SIDE_LEFT
            b.x = O.clamp(-sp.pitch_half_length, bx, sp.pitch_half_length)
            b.y = sp.pitch_half_width if by > 0.0 else -sp.pitch_half_width
            b.vx = b.vy = 0.0
            m.timer = 0
    else:
        m.timer += 1
        if m.mode == PM_AfterGoal:
            if m.timer >= AFTER_GOAL_WAIT:  # kick-off for the side that conceded
                place_formation(m, seed, gid)
                m.mode, m.timer = O.PM_KickOff, 0
                m.side = O.SIDE_RIGHT if m.side == O.SIDE_LEFT else O.SIDE_LEFT
        elif m.timer >= DROP_BALL_TIME:
            m.mode, m.timer = O.PM_PlayOn, 0
    if m.mode != O.PM_PlayOn:
        m.offside = 0
    if not stopped:
        for j, p in enumerate(m.players):
            O.update_stamina(p, sps[j])
        m.cycle += 1
    reward = (goal_l - goal_r) * 10.0 + (O.f32(bx_phys) - O.f32(pbx)) * 0.01
    done = m.step_number >= 2 * half_time
    result = 0 if not done else 1 if m.score_l > m.score_r else 2 if m.score_r > m.score_l else 3
    if done:
        m.mode = O.PM_TimeOver
    return reward, done, result


def collide(m: Match, sp, ball_fixed: bool, model: int = 0):
    """Stadium::collisions as the spec states it: up to 10 rounds; in a round every object collects the positions
    proposed for it and moves to their average; what collided gets vel *= -0.1 once.  Returns who touched the ball."""
    b, pl, n = m.ball, m.players, m.n
    b.collided = False
    touched = [False] * n
    for p in pl:
        p.collided = False
    r, r2 = sp.player_size + sp.ball_size, 2.0 * sp.player_size
    for _ in range(10):
        col = False
        prop = [[0.0, 0.0, 0] for _ in range(n)]
        bsx = bsy = 0.0
        bcnt = 0
        for i, pi in enumerate(pl):
            for j, pj in enumerate(pl):
                if i == j:
                    if ball_fixed:
                        continue
                    dx, dy = b.x - pi.x, b.y - pi.y
                    if dx * dx + dy * dy < r * r:
                        col = b.collided = pi.collided = touched[i] = True
                        nx, ny = O._ball_back_trace(b, pi, r + O.COLLIDE_EPS)
                        bsx, bsy, bcnt = bsx + nx, bsy + ny, bcnt + 1
                        qx, qy = (pi.x, pi.y) if model == 0 else \
                            O._trace_back(pi.x, pi.y, pi.vx, pi.vy, b.x, b.y, r + O.COLLIDE_EPS, -1.0)
                        prop[i][0] += qx
                        prop[i][1] += qy
                        prop[i][2] += 1
                else:
                    ex, ey = pi.x - pj.x, pi.y - pj.y
                    if ex * ex + ey * ey < r2 * r2:
                        col = pi.collided = True
                        if model == 1:  # BACKTRACE: the player backs up along its own velocity
                            qx, qy = O._trace_back(pi.x, pi.y, pi.vx, pi.vy, pj.x, pj.y, r2 + O.COLLIDE_EPS, 1.0 if i < j else -1.0)
                            prop[i][0] += qx
                            prop[i][1] += qy
                            prop[i][2] += 1
                            continue
                        d = math.hypot(ex, ey)
                        ux, uy = ((1.0 if i < j else -1.0), 0.0) if d < 1.0e-10 else (ex / d, ey / d)
                        h = r2 / 2.0 + O.COLLIDE_EPS
                        prop[i][0] += (pi.x + pj.x) / 2.0 + ux * h
                        prop[i][1] += (pi.y + pj.y) / 2.0 + uy * h
                        prop[i][2] += 1
        if bcnt:
            b.x, b.y = bsx / bcnt, bsy / bcnt
        for i, p in enumerate(pl):
            if prop[i][2]:
                p.x, p.y = prop[i][0] / prop[i][2], prop[i][1] / prop[i][2]
        if not col:
            break
    if b.collided:
        b.vx *= -0.1
        b.vy *= -0.1
    for p in pl:
        if p.collided:
            p.vx *= -0.1
            p.vy *= -0.1
    return touched
