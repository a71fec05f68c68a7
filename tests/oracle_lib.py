"""ctypes access to the C oracle (oracle/s2d_oracle.c -> oracle/_build/liboracle_{f64,f32}.so).  Test-side only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

STATE_FIELDS = ("px py vx vy body stamina effort recovery capacity bx by bvx bvy mem_dist mem_ang ep_return "
                "step_number cycle episode flags").split()


# ---- S2DConfig as the oracle sees it: an independent ctypes mirror of include/soccer2d.h, so that a process that only
# ---- times the CPU arm (bench.py --impl reference) never imports the product package or maps libsoccer2d.so.
# ---- tests/test_oracle_c.py checks it field by field against soccer2d_b200._abi.Config.
SP_FIELDS = (
    "pitch_half_length pitch_half_width goal_width goal_post_radius ball_size ball_decay ball_rand ball_speed_max "
    "ball_accel_max player_size player_decay player_rand player_speed_max player_accel_max dash_power_rate inertia_moment "
    "min_dash_power max_dash_power min_dash_angle max_dash_angle dash_angle_step side_dash_rate back_dash_rate "
    "min_power max_power min_moment max_moment kick_power_rate kickable_margin kick_rand "
    "stamina_max stamina_inc_max extra_stamina stamina_capacity recover_init recover_min recover_dec recover_dec_thr "
    "effort_init effort_max effort_min effort_dec effort_dec_thr effort_inc effort_inc_thr "
    "slowness_on_top_for_left_team slowness_on_top_for_right_team").split()


class OracleServerParam(C.Structure):
    _fields_ = [(n, C.c_float) for n in SP_FIELDS] + [("reserved", C.c_float * 5)]


class OracleConfig(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("scenario", C.c_int32), ("num_envs", C.c_int64), ("env_id_offset", C.c_int64),
                ("seed", C.c_uint64), ("device", C.c_int32), ("action_mode", C.c_int32), ("action_space_size", C.c_int32),
                ("max_steps", C.c_int32), ("auto_reset", C.c_int32), ("change_ball_position", C.c_int32),
                ("change_ball_velocity", C.c_int32), ("noise", C.c_int32), ("players_per_side", C.c_int32),
                ("half_time_cycles", C.c_int32), ("kick_actions", C.c_int32), ("collision_model", C.c_int32), ("reserved_i", C.c_int32 * 2),
                ("min_distance_to_ball", C.c_float), ("ball_position_x", C.c_float), ("ball_position_y", C.c_float),
                ("ball_speed", C.c_float), ("ball_direction", C.c_float), ("goto_dist_thr", C.c_float),
                ("reserved_f", C.c_float * 3), ("sp", OracleServerParam)]


ACT_DISCRETE, ACT_CONTINUOUS, ACT_TURNING, ACT_COMMAND = 0, 1, 2, 3


def default_config(num_envs, scenario=0, kind="f64", **kw):
    """the oracle's own S2DConfig defaults (s2do_default_config) with overrides named like the struct fields"""
    cfg = OracleConfig()
    assert lib(kind).s2do_default_config(C.byref(cfg), int(scenario)) == 0
    cfg.num_envs = int(num_envs)
    for k, v in kw.items():
        assert hasattr(cfg, k), k
        setattr(cfg, k, v)
    return cfg


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, capture_output=True)


_libs = {}


def lib(kind: str):
    if kind not in _libs:
        path = os.path.join(ORACLE_DIR, "_build", f"liboracle_{kind}.so")
        src = os.path.join(ORACLE_DIR, "s2d_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            build()
        L = C.CDLL(path)
        L.s2do_create.restype = C.c_int
        L.s2do_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.s2do_destroy.argtypes = [C.c_void_p]
        L.s2do_default_config.restype = C.c_int
        L.s2do_default_config.argtypes = [C.c_void_p, C.c_int]
        L.s2do_set_threads.restype = C.c_int
        L.s2do_set_threads.argtypes = [C.c_int]
        L.s2do_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.s2do_reset_masked.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.s2do_step.argtypes = [C.c_void_p] + [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.s2do_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.s2do_get_state.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.s2do_set_state.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.s2do_obs_dim.argtypes = [C.c_void_p]
        L.s2do_get_state_fg.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.s2do_set_state_fg.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.s2do_get_extra_fg.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.s2do_set_extra_fg.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.s2do_set_ball_fg.argtypes = [C.c_void_p, C.c_int64] + [C.c_double] * 4 + [C.c_int] * 3
        L.s2do_set_player_fg.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_double] * 5
        L.s2do_generate_player_types.restype = C.c_int
        L.s2do_generate_player_types.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
        L.s2do_set_player_types.restype = C.c_int
        L.s2do_set_player_types.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.s2do_set_player_types_per_match.restype = C.c_int
        L.s2do_set_player_types_per_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.s2do_probe_sincos_deg.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.s2do_probe_atan2_deg.restype = C.c_double
        L.s2do_probe_atan2_deg.argtypes = [C.c_double, C.c_double]
        L.s2do_probe_softmax_first.restype = C.c_double
        L.s2do_probe_softmax_first.argtypes = [C.c_double, C.c_double]
        L.s2do_probe_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        _libs[kind] = L
    return _libs[kind]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleSim:
    """N ReachBall episodes stepped on the CPU by the C oracle.  kind = "f64" (truth) or "f32" (the bit-exact
    mirror of the CUDA arithmetic).  `cfg` is a soccer2d_b200._abi.Config (same struct as S2DConfig)."""

    def __init__(self, cfg, kind="f64"):
        self.L = lib(kind)
        self.kind = kind
        self.real = np.float64 if kind == "f64" else np.float32
        assert self.L.s2do_is_f32() == (kind == "f32")
        self.cfg = cfg
        self.n = int(cfg.num_envs)
        self.h = C.c_void_p()
        rc = self.L.s2do_create(C.byref(cfg), C.byref(self.h))
        if rc != 0:
            raise ValueError("oracle rejected the config")
        self.obs_dim = int(self.L.s2do_obs_dim(self.h))
        self.obs = np.zeros((self.n, self.obs_dim), self.real)
        self.term_obs = np.zeros((self.n, self.obs_dim), self.real)
        self.reward = np.zeros(self.n, self.real)
        self.done = np.zeros(self.n, np.uint8)
        self.result = np.zeros(self.n, np.uint8)

    def close(self):
        if self.h is not None:
            self.L.s2do_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def reset(self, mask=None):
        if mask is None:
            self.L.s2do_reset(self.h, _ptr(self.obs))
        else:
            m = np.ascontiguousarray(mask, dtype=np.uint8)
            self.L.s2do_reset_masked(self.h, _ptr(m), _ptr(self.obs))
        return self.obs

    def step(self, actions, k=1):
        a = np.ascontiguousarray(actions)
        self.L.s2do_step(self.h, _ptr(a), int(k), _ptr(self.obs), _ptr(self.reward), _ptr(self.done),
                         _ptr(self.result), _ptr(self.term_obs))
        return self.obs, self.reward, self.done, self.result

    def set_player_types(self, types_array, n, type_of_player):
        """types_array: ctypes array of the product's PlayerType struct (same layout as S2DPlayerType); type_of_player:
        [num_players], or [num_envs][num_players] for an assignment per match"""
        tof = np.ascontiguousarray(np.asarray(type_of_player, dtype=np.uint8))
        setter = self.L.s2do_set_player_types_per_match if tof.ndim == 2 else self.L.s2do_set_player_types
        assert setter(self.h, C.byref(types_array), int(n), _ptr(tof)) == 0

    def stats(self, stats_struct):
        self.L.s2do_stats(self.h, C.byref(stats_struct))
        return stats_struct

    def get_state(self, i=None):
        if i is None:
            return np.stack([self.get_state(j) for j in range(self.n)])
        out = np.zeros(20, np.float64)
        self.L.s2do_get_state(self.h, int(i), _ptr(out))
        return out

    def get_state_fg(self, i=None):
        """FULLGAME: [np*12 + 5 + 12] doubles per env (layout: s2do_get_state_fg)."""
        if i is None:
            return np.stack([self.get_state_fg(j) for j in range(self.n)])
        np_ = 2 * int(self.cfg.players_per_side)
        out = np.zeros(np_ * 12 + 17, np.float64)
        assert self.L.s2do_get_state_fg(self.h, int(i), _ptr(out)) == out.size
        return out

    def get_extra_fg(self, i=None):
        """FULLGAME: [np tackle counters, catch ban of the left keeper, of the right keeper] per env"""
        if i is None:
            return np.stack([self.get_extra_fg(j) for j in range(self.n)])
        out = np.zeros(2 * int(self.cfg.players_per_side) + 2, np.float64)
        assert self.L.s2do_get_extra_fg(self.h, int(i), _ptr(out)) == out.size
        return out

    def set_extra_fg(self, extras):
        ex = np.ascontiguousarray(extras, dtype=np.float64)
        for i in range(self.n):
            self.L.s2do_set_extra_fg(self.h, i, _ptr(ex[i]))

    def set_state_fg(self, states):
        """FULLGAME: overwrite every env's state from an [N, np*12+17] array (layout of get_state_fg)."""
        st = np.ascontiguousarray(states, dtype=np.float64)
        for i in range(self.n):
            self.L.s2do_set_state_fg(self.h, i, _ptr(st[i]))

    def set_state(self, i, vec20):
        v = np.ascontiguousarray(vec20, dtype=np.float64)
        self.L.s2do_set_state(self.h, int(i), _ptr(v))
