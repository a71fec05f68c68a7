"""GPU: the ctypes stub printed in INTEGRATION.md (what a maintainer of the reference would paste into its tree) is
executed as written - with only the library path filled in - and must reproduce the oracle."""
import os
import re
import sys

import numpy as np
import pytest

import helpers as H
import oracle_lib as OL
from soccer2d_b200 import _abi

pytestmark = pytest.mark.gpu


def test_integration_md_stub_runs_and_matches_the_oracle():
    text = open(os.path.join(H.ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    assert len(blocks) >= 2
    code = (blocks[0] + "\n" + blocks[1]).replace('"libsoccer2d.so"', repr(_abi.LIB_PATH))
    sys.path.insert(0, os.path.join(H.ROOT, "tests", "golden", "_shims"))  # `gym` stand-in (gym is not installed)
    try:
        ns = {}
        exec(compile(code, "INTEGRATION.md", "exec"), ns)
    finally:
        sys.path.pop(0)
    env = ns["ReachBallEnv"](use_continuous_action=False, action_space_size=16, max_steps=25)
    cfg = H.make_config(1, "discrete", auto_reset=0, max_steps=25)
    sim = OL.OracleSim(cfg, "f32")
    rng = np.random.default_rng(0)
    assert np.array_equal(env.reset(), sim.reset()[0])
    done = False
    while not done:
        a = int(rng.integers(16))
        obs, reward, done, info = env.step(a)
        sim.step(np.array([[a]], np.uint8))
        assert np.array_equal(obs, sim.obs[0]) and reward == float(sim.reward[0]) and done == bool(sim.done[0])
    assert info["result"] in ("Goal", "Out", "Timeout")
    env.close()
