"""`State` of one env in the shape of the reference's wire schema (idl/service.proto:354-359 State,
:306-349 WorldModel, :181-223 Self, :68-86 Ball, :144-175 Player).

The GPU path never builds protobuf messages; this module turns an `EnvSnapshot` (s2d_export_env) into nested dicts
whose keys are the proto field names, so that code written against the reference's hooks
(`state_to_observation(state: pb2.State)`, soccer_2d_env.py:327-335) or CLSF tooling can be fed with

    from google.protobuf import json_format
    state = json_format.ParseDict(state_dict(snapshot, unum=1), service_pb2.State())

`service_pb2` is the reference's generated module and is not vendored here (protoc is not part of this build).
"""
from __future__ import annotations

import math

from . import _abi

GAME_MODE_NAMES = ("BeforeKickOff", "TimeOver", "PlayOn", "KickOff_", "KickIn_", "FreeKick_", "CornerKick_",
                   "GoalKick_", "AfterGoal_")  # GameModeType, idl/service.proto:267-301
SIDE_NAMES = ("UNKNOWN", "LEFT", "RIGHT")     # Side, idl/service.proto:88-92


def _vec(x, y):
    return {"x": float(x), "y": float(y)}


def _norm_deg(d):
    d = math.fmod(d, 360.0)
    return d - 360.0 if d > 180.0 else d + 360.0 if d < -180.0 else d


def _player(p, ball, ident, kickable_area, type_id=0):
    dx, dy = ball[0] - p.x, ball[1] - p.y
    return {
        "type_id": int(type_id),
        "position": _vec(p.x, p.y), "seen_position": _vec(p.x, p.y), "velocity": _vec(p.vx, p.vy),
        "seen_velocity": _vec(p.vx, p.vy), "id": ident, "side": SIDE_NAMES[p.side], "uniform_number": int(p.uniform_number),
        "is_goalie": p.uniform_number == 1, "body_direction": float(p.body_direction),
        "face_direction": float(p.body_direction), "is_kicking": bool(p.kicked),
        "dist_from_ball": math.hypot(dx, dy), "angle_from_ball": math.degrees(math.atan2(-dy, -dx)) if (dx or dy) else 0.0,
    }


def state_dict(snap: _abi.EnvSnapshot, unum: int = 1, side: int = 1, kickable_area: float = 1.085,
               type_of_player=None) -> dict:
    """proto `State` (as a dict) seen by player `unum` of `side` (1 = LEFT, 2 = RIGHT) in full-state mode: the
    world model is the true state, all *_count fields are 0 (just seen).  `type_of_player[j]`: the heterogeneous type
    of snapshot player j (Player.type_id / Self.type_id, idl/service.proto:174, :216); default 0."""
    players = [snap.players[j] for j in range(snap.num_players)]
    tid = {id(p): (int(type_of_player[j]) if type_of_player is not None else 0) for j, p in enumerate(players)}
    me = next((p for p in players if p.side == side and p.uniform_number == unum), None)
    if me is None:
        raise ValueError(f"no player {unum} on side {side} in this env")
    ball = (snap.ball_x, snap.ball_y, snap.ball_vx, snap.ball_vy)
    dx, dy = ball[0] - me.x, ball[1] - me.y
    dist = math.hypot(dx, dy)
    ang = math.degrees(math.atan2(dy, dx)) if (dx or dy) else 0.0
    mates = [p for p in players if p.side == side and p is not me]
    opps = [p for p in players if p.side != side]
    ident = {id(p): k + 1 for k, p in enumerate(players)}
    our_score, their_score = (snap.left_score, snap.right_score) if side == 1 else (snap.right_score, snap.left_score)
    kick_mates = [p for p in mates if math.hypot(ball[0] - p.x, ball[1] - p.y) <= kickable_area]
    kick_opps = [p for p in opps if math.hypot(ball[0] - p.x, ball[1] - p.y) <= kickable_area]
    wm = {
        "our_team_name": "left" if side == 1 else "right", "their_team_name": "right" if side == 1 else "left",
        "our_side": SIDE_NAMES[side],
        "self": {
            "position": _vec(me.x, me.y), "seen_position": _vec(me.x, me.y), "velocity": _vec(me.vx, me.vy),
            "seen_velocity": _vec(me.vx, me.vy), "id": ident[id(me)], "side": SIDE_NAMES[me.side],
            "uniform_number": int(me.uniform_number), "is_goalie": me.uniform_number == 1,
            "body_direction": float(me.body_direction), "face_direction": float(me.body_direction),
            "is_kicking": bool(me.kicked), "dist_from_ball": dist, "angle_from_ball": _norm_deg(ang + 180.0),
            "stamina": float(me.stamina), "is_kickable": dist <= kickable_area, "recovery": float(me.recovery),
            "stamina_capacity": float(me.stamina_capacity), "effort": float(me.effort), "type_id": tid[id(me)],
        },
        "ball": {
            "position": _vec(ball[0], ball[1]), "relative_position": _vec(dx, dy), "seen_position": _vec(ball[0], ball[1]),
            "velocity": _vec(ball[2], ball[3]), "seen_velocity": _vec(ball[2], ball[3]), "dist_from_self": dist,
            "angle_from_self": ang,
        },
        "teammates": [_player(p, ball, ident[id(p)], kickable_area, tid[id(p)]) for p in mates],
        "opponents": [_player(p, ball, ident[id(p)], kickable_area, tid[id(p)]) for p in opps],
        "our_players_dict": {int(p.uniform_number): _player(p, ball, ident[id(p)], kickable_area, tid[id(p)]) for p in mates + [me]},
        "their_players_dict": {int(p.uniform_number): _player(p, ball, ident[id(p)], kickable_area, tid[id(p)]) for p in opps},
        "our_goalie_uniform_number": 1, "their_goalie_uniform_number": 1 if opps else 0,
        "kickable_teammate_id": ident[id(kick_mates[0])] if kick_mates else 0,
        "kickable_opponent_id": ident[id(kick_opps[0])] if kick_opps else 0,
        "kickable_teammate_existance": bool(kick_mates), "kickable_opponent_existance": bool(kick_opps),
        "cycle": int(snap.cycle), "stoped_cycle": int(snap.stoped_cycle),
        "game_mode_type": GAME_MODE_NAMES[snap.game_mode_type], "game_mode_side": SIDE_NAMES[snap.game_mode_side],
        "left_team_score": int(snap.left_score), "right_team_score": int(snap.right_score),
        "our_team_score": int(our_score), "their_team_score": int(their_score),
        "is_our_set_play": snap.game_mode_type not in (2, 1, 0) and snap.game_mode_side == side,
        "is_their_set_play": snap.game_mode_type not in (2, 1, 0) and snap.game_mode_side not in (0, side),
        "see_time": int(snap.cycle),
    }
    return {"world_model": wm, "full_world_model": wm, "need_preprocess": False}


def trainer_state_dict(snap: _abi.EnvSnapshot) -> dict:
    """What the trainer's State carries on the reference path: the ball and every left player as `teammates`
    (reach_ball_env.py:115-122 reads world_model.teammates[0])."""
    ball = (snap.ball_x, snap.ball_y, snap.ball_vx, snap.ball_vy)
    players = [snap.players[j] for j in range(snap.num_players)]
    wm = {
        "ball": {"position": _vec(ball[0], ball[1]), "velocity": _vec(ball[2], ball[3])},
        "teammates": [_player(p, ball, k + 1, 1.085) for k, p in enumerate(players) if p.side == 1],
        "opponents": [_player(p, ball, k + 1, 1.085) for k, p in enumerate(players) if p.side == 2],
        "cycle": int(snap.cycle), "stoped_cycle": int(snap.stoped_cycle),
        "game_mode_type": GAME_MODE_NAMES[snap.game_mode_type], "game_mode_side": SIDE_NAMES[snap.game_mode_side],
        "left_team_score": int(snap.left_score), "right_team_score": int(snap.right_score),
    }
    return {"world_model": wm}


def player_type_dict(type_id: int, t: dict, sp) -> dict:
    """proto PlayerType (idl/service.proto:1697-1732) of one of the handle's player types: `t` = the fields of
    _abi.PlayerType, `sp` = the env's ServerParam (for the fields rcssserver's default player.conf does not vary and
    for the derived ones: kickable_area, real_speed_max, ...)."""
    decay, dpr, emax = float(t["player_decay"]), float(t["dash_power_rate"]), float(t["effort_max"])
    real_speed_max = min(emax * dpr * sp.max_dash_power / (1.0 - decay), sp.player_speed_max)
    accel = emax * dpr * sp.max_dash_power
    speed, cycles = 0.0, 0
    while cycles < 50 and speed < real_speed_max - 0.01:  # librcsc PlayerType::cyclesToReachMaxSpeed
        speed = speed * decay + accel
        cycles += 1
    return {
        "id": int(type_id), "stamina_inc_max": float(t["stamina_inc_max"]), "player_decay": decay,
        "inertia_moment": float(t["inertia_moment"]), "dash_power_rate": dpr, "player_size": float(sp.player_size),
        "kickable_margin": float(t["kickable_margin"]), "kick_rand": float(t["kick_rand"]),
        "extra_stamina": float(t["extra_stamina"]), "effort_max": emax, "effort_min": float(t["effort_min"]),
        "kick_power_rate": float(t["kick_power_rate"]),
        "kickable_area": float(sp.player_size + sp.ball_size + t["kickable_margin"]),
        "real_speed_max": real_speed_max, "player_speed_max2": float(sp.player_speed_max) ** 2,
        "real_speed_max2": real_speed_max ** 2, "cycles_to_reach_max_speed": cycles,
        "player_speed_max": float(sp.player_speed_max),
    }
