"""Game-log writer: one env's trajectory as an rcssserver game log (`.rcg`, text version 5) so that it can be replayed
in the usual viewers (rcssmonitor / soccerwindow2) - SURVEY.md section 8(f), the debugging / visualisation side of
the path.  Input: the `EnvSnapshot`s of `Soccer2DVecEnv.export_env(i)`, one per cycle.

Layout of a cycle (rcssserver's logger, version 4/5 text format):
    (playmode <t> <name>)                      when the mode changes
    (team <t> <left> <right> <score_l> <score_r>)   when a score changes
    (player_type (id 0)(player_speed_max ..)(stamina_inc_max ..) ...)   header, one per heterogeneous type (optional)
    (show <t> ((b) x y vx vy)
              ((l 1) <type> <state> x y vx vy body neck (v h 180) (s stamina effort recovery capacity)
                     (c kick dash turn catch move tneck view say tackle pointto attention)) ...)
Written from the format's published description; no viewer is available offline to replay the files here.
"""
from __future__ import annotations

PLAYMODE_NAMES = {0: "before_kick_off", 1: "time_over", 2: "play_on", 3: "kick_off", 4: "kick_in", 5: "free_kick",
                  6: "corner_kick", 7: "goal_kick", 8: "after_goal"}
STATE_STAND, STATE_KICK, STATE_GOALIE, STATE_BALL_COLLIDE, STATE_PLAYER_COLLIDE = 0x1, 0x2, 0x8, 0x400, 0x800


def _f(v: float) -> str:
    s = f"{float(v):.4f}".rstrip("0").rstrip(".")
    return "0" if s in ("-0", "") else s


class RcgWriter:
    def __init__(self, path: str, left: str = "left", right: str = "right", player_types: list | None = None,
                 type_of_player=None):
        """player_types: the dicts of `proto_state.player_type_dict` (one `(player_type ...)` header line each, as
        rcssserver logs its heterogeneous types); type_of_player[j]: the type id shown for snapshot player j."""
        self.f = open(path, "w")
        self.left, self.right = left, right
        self.last_mode = None
        self.last_score = None
        self.type_of_player = type_of_player
        self.f.write("ULG5\n")
        for t in player_types or []:
            fields = ("id", "player_speed_max", "stamina_inc_max", "player_decay", "inertia_moment", "dash_power_rate",
                      "player_size", "kickable_margin", "kick_rand", "extra_stamina", "effort_max", "effort_min",
                      "kick_power_rate")
            self.f.write("(player_type " + "".join(f"({k} {_f(t[k]) if k != 'id' else int(t[k])})" for k in fields) + ")\n")

    def _playmode_name(self, snap) -> str:
        name = PLAYMODE_NAMES.get(int(snap.game_mode_type), "play_on")
        if name in ("kick_off", "kick_in", "free_kick", "corner_kick", "goal_kick", "after_goal"):
            name += "_l" if snap.game_mode_side == 1 else "_r"
        return name

    def write(self, snap) -> None:
        """append one cycle (an _abi.EnvSnapshot)"""
        t = int(snap.cycle)
        mode = self._playmode_name(snap)
        if mode != self.last_mode:
            self.f.write(f"(playmode {t} {mode})\n")
            self.last_mode = mode
        score = (int(snap.left_score), int(snap.right_score))
        if score != self.last_score:
            self.f.write(f"(team {t} {self.left} {self.right} {score[0]} {score[1]})\n")
            self.last_score = score
        parts = [f"(show {t} ((b) {_f(snap.ball_x)} {_f(snap.ball_y)} {_f(snap.ball_vx)} {_f(snap.ball_vy)})"]
        for j in range(int(snap.num_players)):
            p = snap.players[j]
            state = STATE_STAND | (STATE_KICK if p.kicked else 0) | (STATE_GOALIE if p.uniform_number == 1 else 0)
            state |= STATE_PLAYER_COLLIDE if p.collided else 0
            side = "l" if p.side == 1 else "r"
            tid = int(self.type_of_player[j]) if self.type_of_player is not None else 0
            parts.append(f" (({side} {int(p.uniform_number)}) {tid} {hex(state)} {_f(p.x)} {_f(p.y)} {_f(p.vx)} {_f(p.vy)} "
                         f"{_f(p.body_direction)} 0 (v h 180) (s {_f(p.stamina)} {_f(p.effort)} {_f(p.recovery)} "
                         f"{_f(p.stamina_capacity)}) (c 0 0 0 0 0 0 0 0 0 0 0))")
        parts.append(")\n")
        self.f.write("".join(parts))

    def close(self) -> None:
        if not self.f.closed:
            self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class RclWriter:
    """The commands of a game as rcssserver's text log (`.rcl`) writes them: one line per player and cycle,
    `<cycle>,<stopped>\tRecv <team>_<unum>: (dash 100 30)`.  Only what the server itself understands can appear in such
    a log - dash / turn / kick (S2D_CMD_DASH / TURN / KICK); the proxy's body actions are lowered before they get here."""
    NAMES = {1: "dash", 2: "turn", 3: "kick"}

    def __init__(self, path: str, left: str = "left", right: str = "right"):
        self.f = open(path, "w")
        self.left, self.right = left, right

    def write(self, cycle: int, commands, players_per_side: int) -> None:
        """commands: [num_players][4] rows {cmd, a, b, c} given in `cycle` (left team first)"""
        for j, c in enumerate(commands):
            name = self.NAMES.get(int(c[0]))
            if name is None:
                continue
            team, unum = (self.left, j + 1) if j < players_per_side else (self.right, j - players_per_side + 1)
            args = f"{_f(c[1])}" if name == "turn" else f"{_f(c[1])} {_f(c[2])}"
            self.f.write(f"{int(cycle)},0\tRecv {team}_{unum}: ({name} {args})\n")

    def close(self) -> None:
        if not self.f.closed:
            self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
