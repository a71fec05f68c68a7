"""Stand-in for `gym` (gym==0.26.2 is pinned by the reference, requirements.txt:3, but is not
installed in this image).  TEST INFRASTRUCTURE ONLY: it exists so that tests/golden/make_golden.py
can import and execute the reference's own sample_environments/reach_ball_env.py.  Only the
three names the reference touches are provided (gym.Env, spaces.Box, spaces.Discrete)."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *args, **kwargs):
        pass
