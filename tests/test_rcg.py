"""CPU-only: the game-log writer produces well-formed version-5 text logs from env snapshots."""
import re

import helpers as H  # noqa: F401
from soccer2d_b200 import _abi
from soccer2d_b200.rcg import RcgWriter


def _snap(cycle, mode, side, score_l):
    s = _abi.EnvSnapshot()
    s.cycle, s.game_mode_type, s.game_mode_side, s.left_score = cycle, mode, side, score_l
    s.ball_x, s.ball_y, s.ball_vx, s.ball_vy = 1.25, -3.0, 0.5, 0.0
    s.num_players = 2
    for j, (x, side_, unum) in enumerate(((-10.0, 1, 1), (10.0, 2, 1))):
        p = s.players[j]
        p.x, p.y, p.body_direction, p.stamina, p.effort, p.recovery, p.stamina_capacity = x, 0.5, 90.0, 7945.0, 1.0, 1.0, 130555.0
        p.side, p.uniform_number, p.kicked = side_, unum, int(j == 0)
    return s


def test_rcg_writer_format(tmp_path):
    path = tmp_path / "game.rcg"
    with RcgWriter(str(path), "b200_l", "b200_r") as w:
        w.write(_snap(1, 3, 1, 0))
        w.write(_snap(2, 2, 0, 0))
        w.write(_snap(3, 2, 0, 1))
    lines = path.read_text().splitlines()
    assert lines[0] == "ULG5"
    assert lines[1] == "(playmode 1 kick_off_l)" and lines[2] == "(team 1 b200_l b200_r 0 0)"
    show = [l for l in lines if l.startswith("(show ")]
    assert len(show) == 3
    pat = re.compile(r"^\(show \d+ \(\(b\) [-\d.]+ [-\d.]+ [-\d.]+ [-\d.]+\)( \(\([lr] \d+\) 0 0x[0-9a-f]+( [-\d.]+){6} \(v h 180\) "
                     r"\(s [-\d.]+ [-\d.]+ [-\d.]+ [-\d.]+\) \(c( 0){11}\)\))+\)$")
    assert all(pat.match(l) for l in show), show[0]
    assert "(playmode 2 play_on)" in lines and "(team 3 b200_l b200_r 1 0)" in lines
    assert "((l 1) 0 0xb -10 0.5" in show[0]  # stand | kick | goalie
    assert all(l.count("(") == l.count(")") for l in lines[1:])


def test_rcg_writer_logs_heterogeneous_player_types(tmp_path):
    import ctypes as C

    from soccer2d_b200.proto_state import player_type_dict
    lib = _abi.load()
    sp = _abi.ServerParam()
    assert lib.s2d_default_server_param(C.byref(sp)) == 0
    types = (_abi.PlayerType * 18)()
    assert lib.s2d_generate_player_types(3, C.byref(sp), types, 18) == 0
    path = tmp_path / "hetero.rcg"
    with RcgWriter(str(path), player_types=[player_type_dict(k, types[k].as_dict(), sp) for k in range(18)],
                   type_of_player=[0, 5]) as w:
        w.write(_snap(1, 2, 0, 0))
    lines = path.read_text().splitlines()
    heads = [l for l in lines if l.startswith("(player_type ")]
    assert len(heads) == 18 and heads[0].startswith("(player_type (id 0)(player_speed_max 1.05)(stamina_inc_max 45)(player_decay 0.4)")
    assert all(l.count("(") == l.count(")") for l in heads)
    show = [l for l in lines if l.startswith("(show ")][0]
    assert "((l 1) 0 0x" in show and "((r 1) 5 0x" in show
