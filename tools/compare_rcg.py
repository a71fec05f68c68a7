#!/usr/bin/env python
"""compare_rcg.py - check the simulator's physics against a real rcssserver game log.

TEST INFRASTRUCTURE (uses the CPU oracle, oracle/s2d_oracle.c, through tests/oracle_lib.py; nothing in the product
imports this).  The rcssserver binary the reference drives (soccer_2d_env.py:356-383, fetched by
scripts/download-rcssserver.sh:30) is not available offline, so the physics restated here has
never met the real thing.  Anybody who has a server can close that gap with this tool:

    rcssserver server::game_logging=1 server::text_logging=1 ...      ->  <stamp>.rcg  (states)  +  <stamp>.rcl  (commands)
    python tools/compare_rcg.py <stamp>.rcg --rcl <stamp>.rcl [--collision-model midpoint|backtrace] [--json report.json]

For every pair of consecutive cycles of the log that are both play_on, the oracle is put into the state of cycle t
(positions, velocities, body directions, stamina / effort / recovery / capacity, as logged), given the commands the server
received in that cycle (dash / turn / kick lines of the .rcl), stepped ONE cycle, and compared with what the log says
about cycle t + 1.  The state is re-synchronised from the log every cycle, so errors do not accumulate: what is reported is the
one-cycle error of every field - count, mean, 99th percentile, maximum, and the share within the log's own resolution (the
.rcg prints four decimals).  Cycles in which the log flags a collision are reported separately and under BOTH collision
models (include/soccer2d.h "Collision models"), which is how to find out which reading of Stadium::collisions the server at
hand implements.  Without an .rcl only command-free laws are checked (ball decay, pos[t+1] - pos[t] = vel[t+1] / decay).

Logs must come from a server run with noise off (server::player_rand=0 ball_rand=0 kick_rand=0 wind_none=1) for the
errors to mean anything; version 4 / 5 text logs (ULG4 / ULG5).  Offline the tool is exercised on a log written by the
product's own writers (soccer2d_b200/rcg.py), see tests/test_compare_rcg.py.
"""
import argparse
import json
import math
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as OL  # noqa: E402

FIELDS = ("ball_x", "ball_y", "ball_vx", "ball_vy", "x", "y", "vx", "vy", "body", "stamina", "effort", "recovery")
RESOLUTION = {"ball_x": 1e-4, "ball_y": 1e-4, "ball_vx": 1e-4, "ball_vy": 1e-4, "x": 1e-4, "y": 1e-4, "vx": 1e-4, "vy": 1e-4,
              "body": 1e-3, "stamina": 1e-2, "effort": 1e-4, "recovery": 1e-4}  # what 4 printed decimals of inputs allow
STATE_BALL_COLLIDE, STATE_PLAYER_COLLIDE = 0x400, 0x800
CMD = {"dash": 1, "turn": 2, "kick": 3}

_SHOW = re.compile(r"^\(show (\d+) ")
_BALL = re.compile(r"\(\(b\) ([-\d.e]+) ([-\d.e]+) ([-\d.e]+) ([-\d.e]+)\)")
_PLAYER = re.compile(r"\(\(([lr]) (\d+)\) (\d+) (0x[0-9a-fA-F]+|\d+) ([-\d.e]+) ([-\d.e]+) ([-\d.e]+) ([-\d.e]+) ([-\d.e]+) ([-\d.e]+)"
                     r"(?: [-\d.e]+ [-\d.e]+)? \(v [hl] [-\d.e]+\) \(s ([-\d.e]+) ([-\d.e]+) ([-\d.e]+)(?: ([-\d.e]+))?\)")
_PARAM = re.compile(r"\((\w+) ([-\d.e]+)\)")


def parse_rcg(path):
    """-> (frames: {cycle: {"ball": [x, y, vx, vy], "players": {(side, unum): dict}, "mode": str}}, server_param overrides)"""
    frames, mode, sp = {}, "before_kick_off", {}
    with open(path) as f:
        for line in f:
            if line.startswith("(playmode "):
                mode = line.split()[2].rstrip(")\n")
            elif line.startswith("(server_param "):
                sp.update({k: float(v) for k, v in _PARAM.findall(line)})
            elif line.startswith("(show "):
                t = int(_SHOW.match(line).group(1))
                b = _BALL.search(line)
                if not b:
                    continue
                players = {}
                for m in _PLAYER.finditer(line):
                    side, unum, _type, state = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4), 0)
                    x, y, vx, vy, body = (float(m.group(k)) for k in range(5, 10))
                    players[(side, unum)] = dict(x=x, y=y, vx=vx, vy=vy, body=body, state=state, stamina=float(m.group(11)),
                                                 effort=float(m.group(12)), recovery=float(m.group(13)),
                                                 capacity=float(m.group(14)) if m.group(14) else -1.0)
                frames[t] = {"ball": [float(b.group(k)) for k in range(1, 5)], "players": players, "mode": mode}
    return frames, sp


def parse_rcl(path, left_name=None, right_name=None):
    """-> {cycle: {(side, unum): [cmd, a, b, 0]}} from lines like  `12,0\\tRecv TeamA_3: (dash 100 30)(turn_neck 0)`;
    the first team named is the left one unless the names are given."""
    out, teams = {}, ([left_name, right_name] if left_name else [])
    pat = re.compile(r"^(\d+),(\d+)\s+Recv (\S+?)_(\d+): (.*)$")
    body = re.compile(r"\((dash|turn|kick) ([-\d.e]+)(?: ([-\d.e]+))?\)")
    with open(path) as f:
        for line in f:
            m = pat.match(line)
            if not m or m.group(3).endswith("Coach"):
                continue
            team = m.group(3)
            if team not in teams and len(teams) < 2:
                teams.append(team)
            if team not in teams:
                continue
            side = "l" if teams.index(team) == 0 else "r"
            c = body.search(m.group(5))
            if c:  # one body command per cycle counts (the first one the server received)
                key = (side, int(m.group(4)))
                out.setdefault(int(m.group(1)), {}).setdefault(key, [CMD[c.group(1)], float(c.group(2)), float(c.group(3) or 0.0), 0.0])
    return out


class Replayer:
    """one oracle FULLGAME env (f64), re-synchronised from the log every cycle"""

    def __init__(self, pps, sp_overrides, collision_model):
        kw = dict(action_mode=OL.ACT_COMMAND, players_per_side=pps, half_time_cycles=10 ** 6, auto_reset=0,
                  collision_model=collision_model)
        self.cfg = OL.default_config(1, 2, **kw)
        for k, v in sp_overrides.items():
            if hasattr(self.cfg.sp, k):
                setattr(self.cfg.sp, k, v)
        self.pps, self.np = pps, 2 * pps
        self.sim = OL.OracleSim(self.cfg, "f64")
        self.sim.reset()

    def slot(self, key):
        side, unum = key
        return (unum - 1) if side == "l" else self.pps + unum - 1

    def step_from(self, frame, commands):
        """state of `frame` + `commands` -> state vector after one cycle (layout of s2do_get_state_fg)"""
        n = self.np
        st = np.zeros(n * 12 + 17)
        P = st[:n * 12].reshape(n, 12)
        for j in range(n):  # absent players wait far outside the pitch, apart from each other
            P[j] = [70.0 + 3.0 * j, 60.0, 0, 0, 0, 8000, 1, 1, 130600, 0, 0, 1 if j < self.pps else 2]
        for key, p in frame["players"].items():
            j = self.slot(key)
            if 0 <= j < n:
                cap = p["capacity"] if p["capacity"] >= 0 else 130600.0
                P[j, :9] = [p["x"], p["y"], p["vx"], p["vy"], p["body"], p["stamina"], p["effort"], p["recovery"], cap]
        k = n * 12
        st[k:k + 4] = frame["ball"]
        st[k + 8] = 2  # play_on
        self.sim.set_state_fg(st[None, :])
        act = np.zeros((1, 1, n, 4), np.float32)
        for key, c in (commands or {}).items():
            j = self.slot(key)
            if 0 <= j < n:
                act[0, 0, j] = c
        self.sim.step(act.reshape(1, -1))
        return self.sim.get_state_fg(0)


def angle_err(a, b):
    d = abs(a - b) % 360.0
    return min(d, 360.0 - d)


def compare(rcg, rcl=None, collision_model="midpoint", models=("midpoint", "backtrace")):
    frames, sp = parse_rcg(rcg)
    cmds = parse_rcl(rcl) if rcl else None
    if not frames:
        raise SystemExit(f"{rcg}: no (show ...) frames found")
    unums = [u for fr in frames.values() for (_s, u) in fr["players"]]
    pps = max(1, min(11, max(unums) if unums else 1))
    model_id = {"midpoint": 0, "backtrace": 1}
    replay = {m: Replayer(pps, sp, model_id[m]) for m in models}
    errs = {m: {"free": {f: [] for f in FIELDS}, "collision": {f: [] for f in FIELDS}} for m in models}
    laws = {"ball_decay": [], "player_integration": []}
    ball_decay = sp.get("ball_decay", 0.94)
    player_decay = sp.get("player_decay", 0.4)
    pairs = 0
    for t in sorted(frames):
        a, b = frames[t], frames.get(t + 1)
        if b is None or a["mode"] != "play_on" or b["mode"] != "play_on":
            continue
        pairs += 1
        collided = any(p["state"] & (STATE_BALL_COLLIDE | STATE_PLAYER_COLLIDE) for p in b["players"].values())
        # command-free laws (always available): MPObject::_inc integrates pos += vel before vel *= decay
        if not collided:
            for key, p1 in b["players"].items():
                p0 = a["players"].get(key)
                if p0:
                    laws["player_integration"].append(math.hypot(p1["x"] - p0["x"] - p1["vx"] / player_decay,
                                                                 p1["y"] - p0["y"] - p1["vy"] / player_decay))
        if cmds is None:
            kickers = [p for p in a["players"].values() if math.hypot(p["x"] - a["ball"][0], p["y"] - a["ball"][1]) < 1.2]
            if not kickers and not collided:
                laws["ball_decay"].append(math.hypot(b["ball"][2] - a["ball"][2] * ball_decay, b["ball"][3] - a["ball"][3] * ball_decay))
            continue
        for m in models:
            rp = replay[m]
            got = rp.step_from(a, cmds.get(t, {}))
            bucket = errs[m]["collision" if collided else "free"]
            k = rp.np * 12
            for f, v, w in zip(FIELDS[:4], got[k:k + 4], b["ball"]):
                bucket[f].append(abs(v - w))
            for key, p1 in b["players"].items():
                j = rp.slot(key)
                if not 0 <= j < rp.np:
                    continue
                g = got[j * 12:j * 12 + 9]
                for f, v in zip(("x", "y", "vx", "vy"), g[:4]):
                    bucket[f].append(abs(v - p1[f]))
                bucket["body"].append(angle_err(g[4], p1["body"]))
                for f, v in zip(("stamina", "effort", "recovery"), g[5:8]):
                    bucket[f].append(abs(v - p1[f]))

    def summary(values, res):
        if not values:
            return None
        v = np.asarray(values)
        return {"n": int(v.size), "mean": float(v.mean()), "p99": float(np.percentile(v, 99)), "max": float(v.max()),
                "within_log_resolution": float((v <= res).mean())}

    report = {"log": os.path.basename(rcg), "frames": len(frames), "play_on_pairs": pairs, "players_per_side": pps,
              "server_param_from_log": sorted(sp), "with_commands": cmds is not None, "selected_model": collision_model,
              "laws": {k: summary(v, 2e-4) for k, v in laws.items()}, "models": {}}
    for m in models:
        report["models"][m] = {kind: {f: summary(errs[m][kind][f], 3 * RESOLUTION[f]) for f in FIELDS} for kind in ("free", "collision")}
    coll = {m: [x for f in ("x", "y", "ball_x", "ball_y") for x in errs[m]["collision"][f]] for m in models}
    if all(coll.values()) and len(models) > 1:
        score = {m: float(np.mean(v)) for m, v in coll.items()}
        report["collision_model_fit"] = {"mean_position_error_on_collision_cycles": score, "best": min(score, key=score.get)}
    return report


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("rcg")
    ap.add_argument("--rcl", help="the server's command log; without it only command-free laws are checked")
    ap.add_argument("--collision-model", default="midpoint", choices=["midpoint", "backtrace"])
    ap.add_argument("--json", help="write the full report here")
    args = ap.parse_args()
    rep = compare(args.rcg, args.rcl, args.collision_model)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rep, f, indent=1)
    print(f"{rep['log']}: {rep['frames']} frames, {rep['play_on_pairs']} play_on cycle pairs, {rep['players_per_side']} a side")
    for k, v in rep["laws"].items():
        if v:
            print(f"  law {k:20s} n={v['n']:7d}  mean {v['mean']:.2e}  p99 {v['p99']:.2e}  max {v['max']:.2e}")
    if rep["with_commands"]:
        for kind in ("free", "collision"):
            print(f"  one-cycle error, {kind} cycles, model {args.collision_model}:")
            for f in FIELDS:
                v = rep["models"][args.collision_model][kind][f]
                if v:
                    print(f"    {f:9s} n={v['n']:7d}  mean {v['mean']:.2e}  p99 {v['p99']:.2e}  max {v['max']:.2e}  "
                          f"within resolution {100 * v['within_log_resolution']:.1f} %")
        if "collision_model_fit" in rep:
            print("  collision cycles, mean position error per model:", rep["collision_model_fit"])
    return 0


if __name__ == "__main__":
    sys.exit(main())
