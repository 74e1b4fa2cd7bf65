"""Host-side restatements of two warp-level algorithms of the CUDA path, checked against plain numpy.

These do not run the kernels (the GPU parity tests do); they pin the *logic* the kernels implement so that a change of
either side has to be made on purpose:

* `knn_quad_kernel` (csrc/knn.cu) selects the K-th smallest squared distance of a target's hits by stepping over distinct
  key values from the previous evaluation's K-th distance and recounting; a tie AT the K-th distance hands the target to
  the tie-breaking kernel.
* `walk_pairs_kernel` (csrc/gravity.cu) keeps (node, target) pairs in a bounded LIFO queue: a round pops
  B = clamp((SOFT - qn) / 7, 1, 32) pairs, each of which pushes at most 8 children one level deeper; cells are only
  expanded into the queue while it holds < 32 pairs.  The queue must never exceed its capacity.
"""
import numpy as np
import pytest

KQ_SEL_STEPS = 8
GP_SOFT, GP_CAP, GP_TMAX, LEVELS = 352, 352 + 160, 16, 21


def select_kth(keys, K, start, steps=KQ_SEL_STEPS):
    """Mirror of the selection in knn_quad_kernel: returns (found, kth, tie)."""
    keys = np.asarray(keys, dtype=np.uint64)
    cur = np.uint64(start)
    c = int((keys <= cur).sum())
    if c < K:
        for _ in range(steps):
            cur = keys[keys > cur].min()          # next distinct value above (exists: c < K <= n)
            c = int((keys <= cur).sum())
            if c >= K:
                return True, int(cur), c > K
    else:
        for _ in range(steps):
            mx = keys[keys <= cur].max()
            ceq = int((keys == mx).sum())
            if c - ceq < K:
                return True, int(mx), c > K
            c -= ceq
            cur = mx - np.uint64(1)
    return False, 0, False


@pytest.mark.parametrize("seed", range(20))
def test_kth_selection_by_stepping(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 97))
    K = 50
    d2 = np.sort(rng.random(n))
    d2[0] = 0.0                                    # the target itself
    keys = np.ascontiguousarray(d2).view(np.uint64)
    rng.shuffle(keys)
    kth_true = np.sort(keys)[K - 1]
    # start a few ranks below / above the true K-th distance, as a slowly changing smoothing length does
    for off in (-5, -1, 0, 1, 4):
        start = np.sort(keys)[min(max(K - 1 + off, 0), n - 1)] + np.uint64(off > 0)
        found, kth, tie = select_kth(keys, K, start)
        assert found and kth == int(kth_true) and not tie
        assert int((keys <= np.uint64(kth)).sum()) == K
    # far off: the step budget runs out and the sort path takes over (found == False), never a wrong answer
    found, kth, tie = select_kth(keys, K, np.sort(keys)[0])
    assert (not found) or kth == int(kth_true)


def test_kth_selection_reports_ties_at_the_boundary_only():
    K = 5
    keys = np.array([0, 10, 20, 20, 30, 40, 40, 50], dtype=np.uint64)     # 5th smallest = 30, unique
    for start in (25, 30, 35, 45):
        assert select_kth(keys, K, start) == (True, 30, False)            # the interior tie (20, 20) is harmless
    keys = np.array([0, 10, 20, 30, 40, 40, 50], dtype=np.uint64)         # 5th and 6th smallest are equal
    for start in (35, 40, 45):
        found, kth, tie = select_kth(keys, K, start)
        assert found and kth == 40 and tie


@pytest.mark.parametrize("p_open", [0.15, 0.5, 1.0])
def test_pair_queue_never_overflows(p_open):
    """Adversarial pushes for a fixed number of rounds: a popped pair opens a cell with (up to) 8 children one level
    deeper with probability p_open (1.0 = every cell above the deepest level has 8 children and is always opened)."""
    rng = np.random.default_rng(int(p_open * 100))
    q = []                                           # depths of the queued pairs (LIFO)
    peak = 0
    rounds = 0
    while rounds < 20000:
        # shared walk: a sparse cell (<= GP_TMAX lanes, <= 8 children) is expanded only while < 32 pairs are queued
        if len(q) < 32:
            depth = int(rng.integers(1, LEVELS))
            q.extend([depth] * (int(rng.integers(1, GP_TMAX + 1)) * int(rng.integers(1, 9))))
            peak = max(peak, len(q))
        while len(q) >= 32 and rounds < 20000:
            rounds += 1
            qn = len(q)
            B = min(max((GP_SOFT - qn) // 7, 1), 32, qn)
            popped = [q.pop() for _ in range(B)]
            for d in popped:
                if d < LEVELS and rng.random() < p_open:
                    q.extend([d + 1] * (8 if p_open == 1.0 else int(rng.integers(1, 9))))
            peak = max(peak, len(q))
            assert len(q) <= GP_CAP, (rounds, len(q))
    assert 32 <= peak <= GP_CAP
