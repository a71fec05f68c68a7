"""Pins oracle/soccer2d_oracle.py (env contract part) against the fixtures produced by executing the
reference's own reach_ball_env.py (tests/golden/make_golden.py)."""
import math

import pytest

from oracle import soccer2d_oracle as O

CMD = {"dash": O.CMD_DASH, "turn": O.CMD_TURN}
RES = {None: 0, "Goal": 1, "Out": 2, "Timeout": 3}


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert O.philox4x32((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert O.philox4x32((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert O.philox4x32((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_decode_discrete(golden):
    for v in golden["decode"]["discrete"]:
        cfg = O.ReachBallConfig(use_continuous_action=False, action_space_size=v["n"])
        cmd, power, d = O.decode_action(cfg, v["a"], 0.5)
        assert (cmd, power, d) == (CMD[v["type"]], v["power"], v["dir"])


def test_decode_continuous(golden):
    cfg = O.ReachBallConfig(use_continuous_action=True)
    for v in golden["decode"]["continuous"]:
        assert O.decode_action(cfg, [v["a"]], 0.5) == (O.CMD_DASH, 100.0, v["dir"])


def test_decode_turning(golden):
    cfg = O.ReachBallConfig(use_continuous_action=True, use_turning=True)
    kinds = set()
    for v in golden["decode"]["turning"]:
        cmd, power, d = O.decode_action(cfg, v["a"], v["u"])
        assert (cmd, power, d) == (CMD[v["type"]], v["power"], v["dir"])
        kinds.add(v["type"])
    assert kinds == {"dash", "turn"}


def test_obs(golden):
    assert len(golden["obs"]) > 200
    for v in golden["obs"]:
        got = O.build_obs(*v["state"])
        for g, w in zip(got, v["obs"]):
            assert g == pytest.approx(w, rel=1e-13, abs=1e-13)


def test_reward_chains(golden):
    seen = set()
    for chain in golden["reward"]:
        cfg = O.ReachBallConfig(**chain["kwargs"])
        md, ma = 0.0, 0.0  # reach_ball_env.py:49-50
        for s in chain["steps"]:
            bx, by, _, _, px, py, body = s["state"]
            done, reward, result, md, ma = O.check_trainer(cfg, md, ma, s["step_number"], bx, by, px, py, body)
            assert done == s["done"] and result == RES[s["result"]]
            assert reward == pytest.approx(s["reward"], rel=1e-12, abs=1e-12)
            assert md == pytest.approx(s["mem_dist"], rel=1e-13) and ma == pytest.approx(s["mem_ang"], rel=1e-13, abs=1e-13)
            seen.add(s["result"])
    assert seen == {None, "Goal", "Out", "Timeout"}


@pytest.mark.parametrize("name", ["default", "dqn_script", "fixed_ball"])
def test_reset_distribution(golden, name):
    """Same supports / moments / acceptance region as reach_ball_env.py:170-218 (stream differs: Philox vs MT)."""
    g = golden["reset"][name]
    cfg = O.ReachBallConfig(seed=99, **g["kwargs"])
    rows = []
    n = 6000
    for e in range(n):
        o = O.ReachBallOracle(cfg, env_id=e)
        rows.append(o.sample_reset())
    cols = list(zip(*rows))  # px, py, body, bx, by, bvx, bvy
    order = [0, 1, 2, 3, 4, 5, 6]
    for k in order[:5]:
        assert all(float(v).is_integer() for v in cols[k])
        assert min(cols[k]) >= g["min"][k] and max(cols[k]) <= g["max"][k]
        if g["std"][k] > 0:
            # uniform integer: mean within 5 sigma of the reference's population mean
            assert abs(sum(cols[k]) / n - g["mean"][k]) < 5 * g["std"][k] / math.sqrt(n) + 5 * g["std"][k] / math.sqrt(g["n"])
            assert min(cols[k]) == g["min"][k] and max(cols[k]) == g["max"][k]
    speed = [math.hypot(a, b) for a, b in zip(cols[5], cols[6])]
    assert max(speed) <= max(g["speed_max"], 3.0) + 1e-6
    assert abs(sum(speed) / n - g["speed_mean"]) < 0.05
    tf = (1.0 - 0.96 ** cfg.max_steps) / (1.0 - 0.96)
    for r, s in zip(rows, speed):
        if s > 0:
            tx, ty = r[3] + r[5] / s * s * tf, r[4] + r[6] / s * s * tf
            assert abs(tx) <= 52.5 + 1e-4 and abs(ty) <= 34.0 + 1e-4
    if name == "dqn_script":
        hist = [0] * 12
        for s in speed:
            hist[min(11, int(s / 0.25))] += 1
        for h, gh in zip(hist, g["speed_hist_0_3_12bins"]):
            p = gh / g["n"]
            assert abs(h / n - p) < 5 * math.sqrt(p * (1 - p) / n) + 0.005
    if name == "fixed_ball":
        assert cols[5][0] == pytest.approx(g["first_rows"][0][5], rel=1e-7)
        assert cols[6][0] == pytest.approx(g["first_rows"][0][6], rel=1e-7)
