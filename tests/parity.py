"""Parity checkers shared by the GPU tests (BASELINE.json north_star rules, written out).

search:  top-k ids identical to the exact fp32 search, except swaps among scores tied
         within TIE_TOL = 1e-3.
embed:   cosine >= 0.999 against the oracle encoder, per sentence.
"""
import numpy as np

TIE_TOL = 1e-3
COS_MIN = 0.999


def check_topk(got_ids, got_scores, ref_ids, ref_scores, exact_scores_of=None, tie_tol=TIE_TOL, score_tol=None):
    """got_* / ref_*: [B, k].  exact_scores_of(b, ids) -> oracle fp32 scores of arbitrary ids.

    Rules: (1) positions whose ids agree must agree; (2) where they differ, the oracle score of
    the returned id must lie within tie_tol of the oracle's score at that position (a swap among
    near-ties, including at the k-th boundary); (3) returned scores are within score_tol of the
    oracle's score for the same id; (4) returned ids are unique.
    Returns the number of positions that differ (all of them legal swaps)."""
    got_ids = np.asarray(got_ids); ref_ids = np.asarray(ref_ids)
    got_scores = np.asarray(got_scores, np.float64); ref_scores = np.asarray(ref_scores, np.float64)
    assert got_ids.shape == ref_ids.shape, (got_ids.shape, ref_ids.shape)
    score_tol = tie_tol if score_tol is None else score_tol
    swaps = 0
    for b in range(got_ids.shape[0]):
        g, r = got_ids[b], ref_ids[b]
        valid = g[g >= 0]
        assert len(set(valid.tolist())) == len(valid), f"query {b}: duplicate ids {g}"
        assert (g >= 0).sum() == (r >= 0).sum(), f"query {b}: {g} vs {r}"
        diff = np.nonzero(g != r)[0]
        if diff.size == 0:
            assert np.all(np.abs(got_scores[b][g >= 0] - ref_scores[b][g >= 0]) <= score_tol), \
                f"query {b}: scores {got_scores[b]} vs {ref_scores[b]}"
            continue
        swaps += int(diff.size)
        if exact_scores_of is not None:
            ex = np.asarray(exact_scores_of(b, g[diff]), np.float64)
        else:
            # the id must at least be in the oracle's list
            pos = {int(i): j for j, i in enumerate(r)}
            assert all(int(i) in pos for i in g[diff]), f"query {b}: ids {g[diff]} not in oracle top-k {r}"
            ex = np.array([ref_scores[b][pos[int(i)]] for i in g[diff]])
        assert np.all(np.abs(ex - ref_scores[b][diff]) <= tie_tol), \
            f"query {b}: non-tie swap at {diff}: got ids {g[diff]} (oracle scores {ex}) vs {r[diff]} ({ref_scores[b][diff]})"
        assert np.all(np.abs(got_scores[b][diff] - ex) <= score_tol), f"query {b}: score mismatch {got_scores[b][diff]} vs {ex}"
    return swaps


def cosine_rows(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return (a * b).sum(1) / (np.linalg.norm(a, axis=1) * np.linalg.norm(b, axis=1))
