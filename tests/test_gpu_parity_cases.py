"""Hand-placed one-player states shared by the host-emulation and the GPU parity tests."""
# px py vx vy body stamina effort recovery capacity | bx by bvx bvy | mem_dist mem_ang ep_ret | step cycle episode
HAND_PLACED_STATES = [
    [0, 0, 0.3, 0.1, 10, 8000, 1, 1, 130600, 0.2, 0.1, -1.0, 0.2, 1, 5, 0, 3, 3, 1],       # ball inside player, moving
    [5, 5, 0, 0, 90, 8000, 1, 1, 130600, 5.1, 5.0, 0, 0, 1, 5, 0, 3, 3, 1],                 # overlapping, ball at rest
    [5, 5, 0, 0, 90, 8000, 1, 1, 130600, 5.0, 5.0, 0, 0, 1, 5, 0, 3, 3, 1],                 # centres coincide
    [-7, 3, 0.5, 0.5, -120, 8000, 1, 1, 130600, -6.8, 3.3, 2.5, -1.0, 1, 5, 0, 3, 3, 1],    # fast ball into player
    [52.3, 0, 0.6, 0, 0, 8000, 1, 1, 130600, 0, 0, 0, 0, 52, 0, 0, 3, 3, 1],                # about to leave over x
    [0, -33.9, 0, -0.6, -90, 8000, 1, 1, 130600, 20, 0, 0, 0, 40, 0, 0, 3, 3, 1],           # about to leave over y
    [10, 10, 0, 0, 0, 30, 0.7, 0.6, 130600, -20, 0, 0, 0, 30, 0, 0, 3, 3, 1],               # nearly exhausted
    [10, 10, 0, 0, 0, 0, 0.6, 0.5, 20, -20, 0, 0, 0, 30, 0, 0, 3, 3, 1],                    # capacity nearly used up
    [10, -10, 0, 0, 45, 2400, 1, 1, 130600, -20, 0, 3.0, 0, 30, 0, 0, 3, 3, 1],             # at the decay thresholds
    [10, -10, 0, 0, 45, 4800, 0.9, 1, 0, -20, 0, 3.0, 0, 30, 0, 0, 3, 3, 1],                # effort_inc threshold, no capacity
]
