"""Size-independent checks of a finished build (test / measurement infrastructure, host side).

`check_pairs` verifies, on sampled ranks j, that (SA[j-1], SA[j]) are in the order the mode defines and
that LCP[j] is the mode's LCP -- by direct comparison on the transformed text.  Together with
`check_positions` (SA is exactly the set of indexed positions) this is the full-size stand-in for the
bit-exact oracle comparison that only fits small inputs.
"""
from __future__ import annotations

import bisect

import numpy as np


def _common_prefix(t: np.ndarray, a: int, b: int, limit: int) -> int:
    """Length of the common prefix of t[a:], t[b:], at most `limit` (vectorised in growing blocks)."""
    done, step = 0, 64
    while done < limit:
        k = min(step, limit - done)
        x, y = t[a + done:a + done + k], t[b + done:b + done + k]
        neq = np.nonzero(x != y)[0]
        if len(neq):
            return done + int(neq[0])
        done += k
        step = min(step * 4, 1 << 22)
    return limit


def check_pairs(text: bytes, sa: np.ndarray, lcp: np.ndarray, ranks, *, seed_mask=None, max_query_len=None,
                n_ranges=()):
    t = np.frombuffer(text, dtype=np.uint8)
    n = len(t)
    starts = [r[0] for r in n_ranges]
    positions = [i for i, c in enumerate(seed_mask) if c == "1"] if seed_mask else None

    def run_end(p):
        k = bisect.bisect_right(starts, p) - 1
        if k >= 0 and n_ranges[k][0] <= p < n_ranges[k][1]:
            return n_ranges[k][1]
        return None

    bad = []
    for j in ranks:
        j = int(j)
        if j == 0:
            if int(lcp[0]) != 0:
                bad.append((j, "lcp[0] != 0"))
            continue
        a, b, l = int(sa[j - 1]), int(sa[j]), int(lcp[j])
        if positions is not None:
            ka = bytes(t[a + o] for o in positions if a + o < n)
            kb = bytes(t[b + o] for o in positions if b + o < n)
            c = 0
            while c < len(ka) and c < len(kb) and ka[c] == kb[c]:
                c += 1
            ok = (ka < kb) or (ka == kb and a > b)
            if not ok or c != l:
                bad.append((j, a, b, l, c))
            continue
        ea, eb = run_end(a), run_end(b)
        if ea is not None and eb is not None and not seed_mask:
            ra, rb = ea - a, eb - b
            want = min(ra, rb)
            if ra != rb:
                ok = (t[a + want] < t[b + want])
            elif t[ea] != t[eb]:
                ok = t[ea] < t[eb]
            else:
                ok = a > b  # tie: larger position first (sufr_builder.rs:701-712)
            if not ok or l != want:
                bad.append((j, a, b, l, want, "n-run"))
            continue
        limit = min(n - a, n - b)
        if max_query_len:
            limit = min(limit, max_query_len)
        c = _common_prefix(t, a, b, limit)
        if c < limit:
            ok = t[a + c] < t[b + c]
        elif max_query_len and c == max_query_len:
            ok = a > b  # our documented tie rule for equal Q-prefixes
        else:
            ok = (n - a) < (n - b)  # a is a proper prefix of b
        if not ok or c != l:
            bad.append((j, a, b, l, c))
    return bad


def check_positions(text: bytes, sa: np.ndarray, *, is_dna=False, allow_ambiguity=False) -> bool:
    """SA holds every indexed position exactly once (sufr_builder.rs:446-449)."""
    t = np.frombuffer(text, dtype=np.uint8)
    if is_dna and not allow_ambiguity:
        keep = (t == ord("$")) | (t == ord("A")) | (t == ord("C")) | (t == ord("G")) | (t == ord("T"))
    else:
        keep = np.ones(len(t), dtype=bool)
    if int(keep.sum()) != len(sa):
        return False
    seen = np.zeros(len(t), dtype=bool)
    seen[sa.astype(np.int64)] = True
    return bool(np.array_equal(seen, keep))
