"""CPU-only: tools/compare_rcg.py (the tool that checks the restated physics against a real rcssserver log) on logs
written by the product's own writers (soccer2d_b200/rcg.py: RcgWriter + RclWriter) from a match the oracle played.
The round trip must close at the log's resolution, and the tool must tell the two collision models apart."""
import os
import sys

import numpy as np
import pytest

import helpers as H
import oracle_lib as OL
from soccer2d_b200 import _abi
from soccer2d_b200.rcg import RclWriter, RcgWriter

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import compare_rcg as CR  # noqa: E402


def snapshot(vec, p, cycle):
    """EnvSnapshot (what Soccer2DVecEnv.export_env returns) from the oracle's FULLGAME state vector"""
    s = _abi.EnvSnapshot()
    k = p * 12
    s.cycle, s.game_mode_type, s.game_mode_side = cycle, int(vec[k + 8]), int(vec[k + 9])
    s.left_score, s.right_score = int(vec[k + 11]), int(vec[k + 12])
    s.ball_x, s.ball_y, s.ball_vx, s.ball_vy = vec[k:k + 4]
    s.ball_collided, s.num_players = int(vec[k + 4]), p
    for j in range(p):
        q, v = s.players[j], vec[j * 12:j * 12 + 12]
        q.x, q.y, q.vx, q.vy, q.body_direction, q.stamina, q.effort, q.recovery, q.stamina_capacity = v[:9]
        q.collided, q.kicked, q.side = int(v[9]), int(v[10]), 1 if j < p // 2 else 2
        q.uniform_number = (j if j < p // 2 else j - p // 2) + 1
    return s


def play_and_log(tmp_path, collision_model, cycles=160, pps=4):
    """a crowded 4 v 4 (everybody dashes at the ball and kicks it: plenty of contacts), server commands only"""
    p = 2 * pps
    cfg = OL.default_config(1, 2, action_mode=OL.ACT_COMMAND, players_per_side=pps, half_time_cycles=10 ** 6, auto_reset=0,
                            collision_model=collision_model, seed=3)
    sim = OL.OracleSim(cfg, "f64")
    sim.reset()
    rng = np.random.default_rng(collision_model)
    rcg, rcl = str(tmp_path / f"m{collision_model}.rcg"), str(tmp_path / f"m{collision_model}.rcl")
    with RcgWriter(rcg, "left", "right") as wg, RclWriter(rcl, "left", "right") as wl:
        st = sim.get_state_fg(0)
        # the log prints four decimals: start from a state that IS what the log says, and keep doing so every cycle
        for t in range(cycles):
            st[:p * 12 + 4] = np.round(st[:p * 12 + 4], 4)
            wg.write(snapshot(st, p, t))
            k = p * 12
            if int(st[k + 8]) != 2:  # the referee stopped play (logged as such): put the ball back and go on
                st[k:k + 4] = [np.round(rng.uniform(-20, 20), 4), np.round(rng.uniform(-15, 15), 4), 0, 0]
                st[k + 8] = 2
            sim.set_state_fg(st[None, :])
            act = np.zeros((1, 1, p, 4), np.float32)
            for j in range(p):
                dx, dy = st[k] - st[j * 12], st[k + 1] - st[j * 12 + 1]
                rel = (np.degrees(np.arctan2(dy, dx)) - st[j * 12 + 4] + 180.0) % 360.0 - 180.0
                if np.hypot(dx, dy) < 0.9 and rng.uniform() < 0.5:
                    act[0, 0, j] = [3, 60.0, float(np.round(rng.uniform(-90, 90))), 0]
                elif abs(rel) > 25.0:
                    act[0, 0, j] = [2, float(np.round(rel)), 0, 0]
                else:
                    act[0, 0, j] = [1, float(np.round(rng.uniform(40, 100))), 0, 0]
            wl.write(t, act[0, 0], pps)
            sim.step(act.reshape(1, -1))
            st = sim.get_state_fg(0)
        st[:p * 12 + 4] = np.round(st[:p * 12 + 4], 4)
        wg.write(snapshot(st, p, cycles))
    return rcg, rcl


@pytest.mark.parametrize("model", ["midpoint", "backtrace"])
def test_round_trip_closes_at_log_resolution_and_identifies_the_collision_model(tmp_path, model):
    rcg, rcl = play_and_log(tmp_path, {"midpoint": 0, "backtrace": 1}[model])
    frames, _ = CR.parse_rcg(rcg)
    cmds = CR.parse_rcl(rcl)
    assert len(frames) == 161 and len(cmds) >= 150 and all(len(v) == 8 for v in cmds.values())
    rep = CR.compare(rcg, rcl, model)
    assert rep["with_commands"] and rep["play_on_pairs"] >= 100 and rep["players_per_side"] == 4
    free = rep["models"][model]["free"]
    for f in ("x", "y", "vx", "vy", "ball_x", "ball_y", "ball_vx", "ball_vy", "effort", "recovery"):
        assert free[f]["max"] < 2.5e-4, (f, free[f])  # the log's own rounding (half a unit in the 4th decimal, in and out)
    assert free["body"]["max"] < 1e-3 and free["stamina"]["max"] < 1e-2
    assert rep["laws"]["player_integration"]["max"] < 5e-4
    # cycles the log flags as collisions: the model that wrote the log fits, the other one does not
    fit = rep["collision_model_fit"]
    assert rep["models"][model]["collision"]["x"]["n"] >= 16, "the scripted match should produce contacts"
    assert fit["best"] == model, fit
    other = "backtrace" if model == "midpoint" else "midpoint"
    assert fit["mean_position_error_on_collision_cycles"][model] < 2e-4 < fit["mean_position_error_on_collision_cycles"][other]


def test_without_a_command_log_only_the_command_free_laws_are_checked(tmp_path):
    rcg, _ = play_and_log(tmp_path, 0, cycles=60)
    rep = CR.compare(rcg)
    assert not rep["with_commands"] and rep["laws"]["player_integration"]["n"] > 100
    assert rep["laws"]["player_integration"]["max"] < 5e-4
    assert all(v is None for v in rep["models"]["midpoint"]["free"].values())


def test_cli_prints_a_report(tmp_path, capsys):
    rcg, rcl = play_and_log(tmp_path, 0, cycles=40)
    sys.argv = ["compare_rcg.py", rcg, "--rcl", rcl, "--json", str(tmp_path / "r.json")]
    assert CR.main() == 0
    out = capsys.readouterr().out
    assert "play_on cycle pairs" in out and "one-cycle error, free cycles" in out and (tmp_path / "r.json").exists()
