"""CPU-only checks of the drop-in boundary: the shared library loads, exports every symbol include/soccer2d.h
declares, agrees with the header on struct sizes and defaults, and FAILS LOUDLY without a GPU (no CPU path)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import helpers as H
from soccer2d_b200 import _abi

HEADER = os.path.join(H.ROOT, "include", "soccer2d.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2d_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared_functions()
    assert len(names) >= 20
    lib = C.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/soccer2d.h but not exported"
    assert sorted(_abi.SIGNATURES) == names  # the Python binding covers exactly the declared surface


def test_abi_version_and_struct_sizes_match_the_header(tmp_path):
    lib = _abi.load()
    assert lib.s2d_abi_version() == _abi.ABI_VERSION
    # ask the C compiler for the truth
    prog = tmp_path / "sizes.c"
    prog.write_text('#include <stdio.h>\n#include "soccer2d.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %d\\n",'
                    "sizeof(S2DServerParam),sizeof(S2DConfig),sizeof(S2DBuffers),sizeof(S2DStats),"
                    "sizeof(S2DPlayerSnapshot),sizeof(S2DEnvSnapshot),sizeof(S2DPlayerType),sizeof(S2DMlpPolicy),sizeof(S2DTrajectory),S2D_ABI_VERSION);return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-I", os.path.dirname(HEADER), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(t) for t in (_abi.ServerParam, _abi.Config, _abi.Buffers, _abi.Stats, _abi.PlayerSnapshot,
                                  _abi.EnvSnapshot, _abi.PlayerType, _abi.MlpPolicy, _abi.Trajectory)] + [_abi.ABI_VERSION]
    assert got == want


def test_defaults_are_the_reference_defaults(golden):
    lib = _abi.load()
    cfg = _abi.Config()
    assert lib.s2d_default_config(C.byref(cfg), _abi.SCENARIO_REACHBALL) == 0
    d = golden["spaces"]["envs"][0]["defaults"]  # produced by the reference's ReachBallEnv.__init__
    assert cfg.max_steps == d["max_steps"] and cfg.action_space_size == d["action_space_size"]
    assert cfg.min_distance_to_ball == d["min_distance_to_ball"]
    assert bool(cfg.change_ball_position) == d["change_ball_position"]
    assert bool(cfg.change_ball_velocity) == d["change_ball_velocity"]
    assert (cfg.action_mode == _abi.ACT_CONTINUOUS) == (d["use_continuous_action"] and not d["use_turning"])
    assert cfg.struct_size == C.sizeof(_abi.Config)
    assert lib.s2d_obs_dim(C.byref(cfg)) == golden["spaces"]["envs"][0]["observation_space"]["shape"][0]
    assert lib.s2d_num_players(C.byref(cfg)) == 1
    cfg.num_envs = 1000
    assert lib.s2d_state_bytes(C.byref(cfg)) == 80 * 1000
    assert lib.s2d_action_bytes(C.byref(cfg)) == 4 * 1000
    # rcssserver defaults, names = proto ServerParam
    sp = cfg.sp
    assert (sp.player_decay, sp.ball_decay, sp.dash_power_rate) == pytest.approx((0.4, 0.94, 0.006))
    assert (sp.stamina_max, sp.stamina_inc_max, sp.kickable_margin) == pytest.approx((8000, 45, 0.7))


def test_invalid_configs_are_rejected_with_a_message():
    lib = _abi.load()
    h = C.c_void_p()
    cfg = H.make_config(4)
    cfg.struct_size = 12
    assert lib.s2d_create(C.byref(cfg), C.byref(h)) == _abi.S2D_ERR_INVALID
    assert b"struct_size" in lib.s2d_last_error(None)
    cfg = H.make_config(0)
    assert lib.s2d_create(C.byref(cfg), C.byref(h)) == _abi.S2D_ERR_INVALID
    cfg = H.make_config(4, action_space_size=300)
    assert lib.s2d_create(C.byref(cfg), C.byref(h)) == _abi.S2D_ERR_INVALID
    assert lib.s2d_create(None, C.byref(h)) == _abi.S2D_ERR_INVALID
    assert lib.s2d_step(None, 1, None) == _abi.S2D_ERR_INVALID
    assert lib.s2d_error_string(_abi.S2D_ERR_NO_DEVICE).startswith(b"no usable CUDA device")


def test_no_gpu_means_error_not_fallback():
    """In a container without a CUDA device the product path must refuse to run."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = _abi.load()
    h = C.c_void_p()
    cfg = H.make_config(4)
    assert lib.s2d_create(C.byref(cfg), C.byref(h)) == _abi.S2D_ERR_NO_DEVICE
    assert not h.value
    from soccer2d_b200 import Soccer2DError, Soccer2DVecEnv
    with pytest.raises(Soccer2DError):
        Soccer2DVecEnv(4)
    from sample_environments.environment_factory import EnvironmentFactory
    with pytest.raises(Soccer2DError):
        EnvironmentFactory().create("ReachBall", None, None, "/tmp")
    with pytest.raises(ValueError, match="Environment ReachCenter not found."):
        EnvironmentFactory().create("ReachCenter", None, None, "/tmp")


def test_product_never_imports_the_oracle():
    pkg = H.PKG
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.lower() or f == "s2d_math.cuh", f  # s2d_math.cuh mentions the test oracle in a comment
                assert "import oracle" not in text and "from oracle" not in text and "liboracle" not in text


def test_player_types_are_rcssserver_hetero_draws_and_match_the_oracle():
    """s2d_generate_player_types is host code: type 0 is the default player, the others trade decay against inertia,
    dash power against stamina income, kickable margin against kick noise and extra stamina against effort, and can
    sustain a speed within 0.1 below player_speed_max.  The oracle's float build draws the very same table."""
    import oracle_lib as OL
    lib = _abi.load()
    sp = _abi.ServerParam()
    assert lib.s2d_default_server_param(C.byref(sp)) == 0
    for seed in (0, 7, 123456789):
        mine = (_abi.PlayerType * 18)()
        assert lib.s2d_generate_player_types(seed, C.byref(sp), mine, 18) == 0
        theirs = (_abi.PlayerType * 18)()
        assert OL.lib("f32").s2do_generate_player_types(seed, C.byref(sp), C.byref(theirs), 18) == 0
        assert bytes(mine) == bytes(theirs)
        truth = (_abi.PlayerType * 18)()
        assert OL.lib("f64").s2do_generate_player_types(seed, C.byref(sp), C.byref(truth), 18) == 0
        t0 = mine[0].as_dict()
        assert t0 == pytest.approx({k: getattr(sp, k) for k in t0})
        distinct = set()
        for k in range(1, 18):
            t, d = mine[k].as_dict(), truth[k].as_dict()
            assert t == pytest.approx(d, rel=2e-6, abs=1e-7)
            distinct.add(round(t["player_decay"], 6))
            assert 0.3 <= t["player_decay"] <= 0.5 and 0.6 <= t["kickable_margin"] <= 0.8
            assert 0.0048 - 1e-7 <= t["dash_power_rate"] <= 0.0068 + 1e-7 and 50.0 <= t["extra_stamina"] <= 100.0
            assert t["inertia_moment"] == pytest.approx(5.0 + (t["player_decay"] - 0.4) * 25.0, abs=1e-4)
            assert t["stamina_inc_max"] == pytest.approx(45.0 - (t["dash_power_rate"] - 0.006) * 6000.0, abs=1e-3)
            assert t["kick_rand"] == pytest.approx(0.1 + (t["kickable_margin"] - 0.7), abs=1e-6)
            assert t["effort_max"] == pytest.approx(1.0 - (t["extra_stamina"] - 50.0) * 0.004, abs=1e-5)
            assert t["effort_min"] == pytest.approx(0.6 - (t["extra_stamina"] - 50.0) * 0.004, abs=1e-5)
            rsm = t["effort_max"] * t["dash_power_rate"] * 100.0 / (1.0 - t["player_decay"])
            assert 0.95 - 1e-6 < rsm < 1.05 + 1e-6
        assert len(distinct) >= 15
    assert lib.s2d_generate_player_types(0, C.byref(sp), mine, 19) == _abi.S2D_ERR_INVALID


def test_python_constants_equal_the_header_defines():
    """every S2D_<GROUP>_<NAME> the Python host mirrors has the header's value (commands, action modes, scenarios,
    results, flags, error codes), and the oracle's Python constants agree for the command vocabulary"""
    import re
    from oracle import soccer2d_oracle as O
    text = open(HEADER).read()
    defines = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+S2D_(\w+)\s+\(?(-?(?:0x[0-9a-fA-F]+|\d+))u?\)?", text)}
    checked = 0
    for name, value in defines.items():
        for attr in (name, "S2D_" + name):
            if hasattr(_abi, attr) and isinstance(getattr(_abi, attr), int):
                assert getattr(_abi, attr) == value, (attr, getattr(_abi, attr), value)
                checked += 1
    assert checked >= 30 and defines["CMD_INTERCEPT"] == 10 and defines["ABI_VERSION"] == _abi.ABI_VERSION
    for name in ("NONE", "DASH", "TURN", "KICK", "GOTO", "TURN_TO_POINT", "TURN_TO_BALL", "TURN_TO_ANGLE", "KICK_ONE_STEP",
                 "STOP_BALL", "INTERCEPT"):
        assert getattr(O, "CMD_" + name) == defines["CMD_" + name]
