"""CPU-only: the fragment bookkeeping of csrc/s2d_rollout.cuh, emulated lane by lane in numpy.

The fused policy kernel never shuffles between layers: the accumulator fragment of one `mma.sync.m16n8k8` layer is fed
to the next layer as its A fragment, which is only correct because the weight rows are paired (2t, 2t + 1) instead of
the instruction's native (t, t + 4).  This test restates that argument executable: a warp of 32 lanes, every fragment
exactly as the PTX ISA lays it out, three layers chained the way `mlp_forward_tile` chains them, against a plain matrix
product.  (It checks the index algebra the kernel is written from - the kernel itself is checked on the GPU against
torch, tests/test_gpu_rollout.py.)"""
import numpy as np


def mma_m16n8k8(d, a, b):
    """d[lane][4] += A(16x8) @ B(8x8) with the fragments of mma.sync.aligned.m16n8k8.row.col (g = lane / 4, t = lane % 4):
    a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (k = t, n = g) b1 (k = t+4, n = g);
    d0 (g, 2t) d1 (g, 2t+1) d2 (g+8, 2t) d3 (g+8, 2t+1)"""
    A = np.zeros((16, 8))
    B = np.zeros((8, 8))
    for lane in range(32):
        g, t = lane // 4, lane % 4
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[lane]
        B[t, g], B[t + 4, g] = b[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane // 4, lane % 4
        d[lane] += [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]


def forward_tile(obs16, W1, b1, W2, b2, W3, b3):
    """Q[16][16] of 16 observations (padded to 16 features) with the kernel's fragment choreography"""
    lanes = range(32)
    gt = [(lane // 4, lane % 4) for lane in lanes]
    # layer 1: A from the staged observations in the native order, B = {W1[n][8k + t], W1[n][8k + t + 4]}
    h1 = [np.array([[b1[8 * j + 2 * t], b1[8 * j + 2 * t + 1]] * 2 for (g, t) in gt], dtype=float) for j in range(8)]
    for k in range(2):
        a = [[obs16[g, 8 * k + t], obs16[g + 8, 8 * k + t], obs16[g, 8 * k + t + 4], obs16[g + 8, 8 * k + t + 4]] for (g, t) in gt]
        for j in range(8):
            b = [[W1[8 * j + g, 8 * k + t], W1[8 * j + g, 8 * k + t + 4]] for (g, t) in gt]
            mma_m16n8k8(h1[j], a, b)
    # layers 2 and 3: the accumulator (g, 2t) (g, 2t+1) (g+8, 2t) (g+8, 2t+1) of column tile k IS the A fragment
    # (g, .) (g+8, .) (g, .) (g+8, .) of k-step k when B = {W[n][8k + 2t], W[n][8k + 2t + 1]}
    def chained(prev, W, bias, n_tiles):
        out = [np.array([[bias[8 * j + 2 * t], bias[8 * j + 2 * t + 1]] * 2 for (g, t) in gt], dtype=float) for j in range(n_tiles)]
        for k in range(8):
            relu = np.maximum(prev[k], 0.0)
            a = [[relu[lane][0], relu[lane][2], relu[lane][1], relu[lane][3]] for lane in lanes]
            for j in range(n_tiles):
                b = [[W[8 * j + g, 8 * k + 2 * t], W[8 * j + g, 8 * k + 2 * t + 1]] for (g, t) in gt]
                mma_m16n8k8(out[j], a, b)
        return out

    h2 = chained(h1, W2, b2, 8)
    q = chained(h2, W3, b3, 2)
    Q = np.zeros((16, 16))
    for lane, (g, t) in enumerate(gt):
        for j in range(2):
            Q[g, 8 * j + 2 * t], Q[g, 8 * j + 2 * t + 1], Q[g + 8, 8 * j + 2 * t], Q[g + 8, 8 * j + 2 * t + 1] = q[j][lane]
    return Q


def test_chained_accumulator_fragments_compute_the_mlp():
    rng = np.random.default_rng(0)
    obs = np.zeros((16, 16))
    obs[:, :10] = rng.uniform(-1, 1, (16, 10))
    W1 = np.zeros((64, 16))
    W1[:, :10] = rng.normal(size=(64, 10))
    b1, W2, b2 = rng.normal(size=64), rng.normal(size=(64, 64)) / 8, rng.normal(size=64)
    W3, b3 = rng.normal(size=(16, 64)) / 8, rng.normal(size=16)
    want = np.maximum(np.maximum(obs @ W1.T + b1, 0.0) @ W2.T + b2, 0.0) @ W3.T + b3
    got = forward_tile(obs, W1, b1, W2, b2, W3, b3)
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_quad_argmax_keeps_the_lowest_index_on_ties():
    """the kernel's argmax: every lane first over its own four values in ascending action order (strict >), then two
    xor-shuffle steps inside the quad that prefer the larger value and, on equal values, the lower index"""
    rng = np.random.default_rng(1)
    for trial in range(200):
        q = rng.integers(0, 4, size=16).astype(float)  # many ties
        best = {}
        for t in range(4):
            cand = [(q[2 * t], 2 * t), (q[2 * t + 1], 2 * t + 1), (q[8 + 2 * t], 8 + 2 * t), (q[8 + 2 * t + 1], 8 + 2 * t + 1)]
            v, a = cand[0]
            for cv, ca in cand[1:]:
                if cv > v:
                    v, a = cv, ca
            best[t] = (v, a)
        for m in (1, 2):
            nxt = {}
            for t in range(4):
                (v, a), (ov, oa) = best[t], best[t ^ m]
                nxt[t] = (ov, oa) if ov > v or (ov == v and oa < a) else (v, a)
            best = nxt
        assert all(best[t][1] == int(np.argmax(q)) for t in range(4))


def mma_m16n8k16(d, a, b):
    """bf16 form, values kept exact here: a0 {(g, 2t), (g, 2t+1)} a1 {(g+8, ..)} a2 {(g, 2t+8), (g, 2t+9)} a3 {(g+8, ..)};
    b0 {(k = 2t, n = g), (2t+1, g)} b1 {(2t+8, g), (2t+9, g)}; d as m16n8k8"""
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane // 4, lane % 4
        (A[g, 2 * t], A[g, 2 * t + 1]), (A[g + 8, 2 * t], A[g + 8, 2 * t + 1]) = a[lane][0], a[lane][1]
        (A[g, 2 * t + 8], A[g, 2 * t + 9]), (A[g + 8, 2 * t + 8], A[g + 8, 2 * t + 9]) = a[lane][2], a[lane][3]
        (B[2 * t, g], B[2 * t + 1, g]), (B[2 * t + 8, g], B[2 * t + 9, g]) = b[lane]
    D = A @ B
    for lane in range(32):
        g, t = lane // 4, lane % 4
        d[lane] += [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]


def test_bf16_form_chains_two_column_tiles_per_k_step():
    """mlp_forward_tile_bf16: two consecutive accumulator column tiles ARE the A fragment of the next k16 step, in order"""
    rng = np.random.default_rng(2)
    obs = np.zeros((16, 16))
    obs[:, :10] = rng.uniform(-1, 1, (16, 10))
    W1 = np.zeros((64, 16))
    W1[:, :10] = rng.normal(size=(64, 10))
    b1, W2, b2 = rng.normal(size=64), rng.normal(size=(64, 64)) / 8, rng.normal(size=64)
    W3, b3 = rng.normal(size=(24, 64)) / 8, rng.normal(size=24)
    gt = [(lane // 4, lane % 4) for lane in range(32)]

    def bias(bv, j):
        return np.array([[bv[8 * j + 2 * t], bv[8 * j + 2 * t + 1]] * 2 for (g, t) in gt], dtype=float)

    def bfrag(W, k, j):
        return [((W[8 * j + g, 16 * k + 2 * t], W[8 * j + g, 16 * k + 2 * t + 1]),
                 (W[8 * j + g, 16 * k + 2 * t + 8], W[8 * j + g, 16 * k + 2 * t + 9])) for (g, t) in gt]

    h1 = [bias(b1, j) for j in range(8)]
    a = [((obs[g, 2 * t], obs[g, 2 * t + 1]), (obs[g + 8, 2 * t], obs[g + 8, 2 * t + 1]),
          (obs[g, 2 * t + 8], obs[g, 2 * t + 9]), (obs[g + 8, 2 * t + 8], obs[g + 8, 2 * t + 9])) for (g, t) in gt]
    for j in range(8):
        mma_m16n8k16(h1[j], a, bfrag(W1, 0, j))

    def chained(prev, W, bv, n_tiles):
        out = [bias(bv, j) for j in range(n_tiles)]
        for k in range(4):
            lo, hi = np.maximum(prev[2 * k], 0.0), np.maximum(prev[2 * k + 1], 0.0)
            a = [((lo[l][0], lo[l][1]), (lo[l][2], lo[l][3]), (hi[l][0], hi[l][1]), (hi[l][2], hi[l][3])) for l in range(32)]
            for j in range(n_tiles):
                mma_m16n8k16(out[j], a, bfrag(W, k, j))
        return out

    q = chained(chained(h1, W2, b2, 8), W3, b3, 3)
    Q = np.zeros((16, 24))
    for lane, (g, t) in enumerate(gt):
        for j in range(3):
            Q[g, 8 * j + 2 * t], Q[g, 8 * j + 2 * t + 1], Q[g + 8, 8 * j + 2 * t], Q[g + 8, 8 * j + 2 * t + 1] = q[j][lane]
    want = np.maximum(np.maximum(obs @ W1.T + b1, 0.0) @ W2.T + b2, 0.0) @ W3.T + b3
    assert np.allclose(Q, want, rtol=1e-12, atol=1e-12)
