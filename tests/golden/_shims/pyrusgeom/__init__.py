"""Stand-in for pyrusgeom==0.1.2 (requirements.txt:6; not installed in this image).
TEST INFRASTRUCTURE ONLY.  Restates the handful of Vector2D / AngleDeg operations that
reach_ball_env.py:89-96,119-124,133,142,206-216 calls.  Decisions that cannot be checked
offline are marked (*) in geom_2d.py."""
