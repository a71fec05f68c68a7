"""Light stand-ins for the handful of `service_pb2` messages the env path touches (idl/service.proto), for hook-style
scenario classes when the reference's generated module is not installed:

    from soccer2d_b200 import pb2_lite as pb2
    pb2.PlayerAction(dash=pb2.Dash(power=100, relative_direction=30.0))
    pb2.TrainerAction(do_move_ball=pb2.DoMoveBall(position=pb2.RpcVector2D(x=0, y=0), velocity=pb2.RpcVector2D(x=1, y=0)))

Same field names, same `WhichOneof("action")`; unset scalar fields read as 0 / False / "" and unset sub-messages as
empty messages, like protobuf.  `View` wraps the dicts of soccer2d_b200.proto_state the same way, so that
`state.world_model.ball.position.x` works on a State built without protobuf.
"""
from __future__ import annotations


class _Message:
    _oneof: tuple = ()

    def __init__(self, **fields):
        for k, v in fields.items():
            object.__setattr__(self, k, v)

    def __getattr__(self, name):  # unset field
        if name.startswith("_"):
            raise AttributeError(name)
        return _Empty()

    def WhichOneof(self, group):  # noqa: N802 - protobuf's name
        for name in self._oneof:
            if name in self.__dict__:
                return name
        return None

    def __repr__(self):
        return f"{type(self).__name__}({', '.join(f'{k}={v!r}' for k, v in self.__dict__.items())})"


class _Empty(_Message):
    """an unset field: behaves as 0 / False / "" and as an empty message / list"""

    def __bool__(self):
        return False

    def __float__(self):
        return 0.0

    def __int__(self):
        return 0

    def __index__(self):
        return 0

    def __len__(self):
        return 0

    def __iter__(self):
        return iter(())

    def __eq__(self, other):
        return other in (0, 0.0, False, "", None) or isinstance(other, _Empty)

    def __hash__(self):
        return 0


def _message(name, oneof=()):
    return type(name, (_Message,), {"_oneof": tuple(oneof)})


RpcVector2D = _message("RpcVector2D")
Dash = _message("Dash")                      # power, relative_direction       (idl/service.proto:380-383)
Turn = _message("Turn")                      # relative_direction              (:390-392)
Kick = _message("Kick")                      # power, relative_direction       (:394-397)
Body_GoToPoint = _message("Body_GoToPoint")  # target_point, distance_threshold, max_dash_power (:684-688)
Body_HoldBall = _message("Body_HoldBall")    # (:748-752)
Body_Intercept = _message("Body_Intercept")      # save_recovery, face_point (:742-745)
Body_KickOneStep = _message("Body_KickOneStep")  # target_point, first_speed, force_mode (:747-751)
Body_SmartKick = _message("Body_SmartKick")      # target_point, first_speed, first_speed_threshold, max_steps (:690-695)
Body_StopBall = _message("Body_StopBall")        # (:753-754)
Body_TurnToAngle = _message("Body_TurnToAngle")  # angle (:769-771)
Body_TurnToBall = _message("Body_TurnToBall")    # cycle (:773-775)
Body_TurnToPoint = _message("Body_TurnToPoint")  # target_point, cycle (:777-780)
PlayerAction = _message("PlayerAction", oneof=("dash", "turn", "kick", "body_go_to_point", "body_hold_ball",
                                               "body_kick_one_step", "body_stop_ball", "body_turn_to_angle",
                                               "body_turn_to_ball", "body_turn_to_point", "body_intercept", "body_smart_kick"))
DoMoveBall = _message("DoMoveBall")          # position, velocity              (:1395-1398)
DoMovePlayer = _message("DoMovePlayer")      # our_side, uniform_number, position, body_direction (:1400-1405)
DoRecover = _message("DoRecover")            # (:1407)
DoChangeMode = _message("DoChangeMode")      # game_mode_type, side            (:1409-1412)
DoKickOff = _message("DoKickOff")
TrainerAction = _message("TrainerAction", oneof=("do_kick_off", "do_move_ball", "do_move_player", "do_recover",
                                                 "do_change_mode", "do_change_player_type"))


class GameModeType:  # idl/service.proto:267-301 (the values the path uses)
    BeforeKickOff, TimeOver, PlayOn, KickOff_, KickIn_, FreeKick_, CornerKick_, GoalKick_, AfterGoal_ = range(9)


class Side:  # idl/service.proto:88-92
    UNKNOWN, LEFT, RIGHT = 0, 1, 2


class View:
    """attribute access over the nested dicts of proto_state.state_dict (enum names become their numbers)"""
    _ENUMS = {"game_mode_type": ("BeforeKickOff", "TimeOver", "PlayOn", "KickOff_", "KickIn_", "FreeKick_", "CornerKick_",
                                 "GoalKick_", "AfterGoal_"),
              "side": ("UNKNOWN", "LEFT", "RIGHT"), "our_side": ("UNKNOWN", "LEFT", "RIGHT"),
              "game_mode_side": ("UNKNOWN", "LEFT", "RIGHT")}

    def __init__(self, d: dict):
        object.__setattr__(self, "_d", d)

    def __getattr__(self, name):
        d = object.__getattribute__(self, "_d")
        if name not in d:
            return _Empty()
        v = d[name]
        if isinstance(v, dict):
            return View(v) if all(isinstance(k, str) for k in v) else {k: View(x) if isinstance(x, dict) else x for k, x in v.items()}
        if isinstance(v, list):
            return [View(x) if isinstance(x, dict) else x for x in v]
        if isinstance(v, str) and name in self._ENUMS and v in self._ENUMS[name]:
            return self._ENUMS[name].index(v)
        return v

    def HasField(self, name):  # noqa: N802
        return name in object.__getattribute__(self, "_d")
