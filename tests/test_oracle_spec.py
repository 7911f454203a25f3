"""Cross-checks the literal restatement (oracle) against the closed-form semantics
(tests/oracle.py::spec_build) on seeded random inputs -- the modes whose result is a pure
function of the input (full sort, seed mask, MQL without ties)."""
import random

import numpy as np
import pytest

import oracle as O


def rand_text(rng, n, alphabet, repeat_p=0.0):
    out = bytearray()
    while len(out) < n:
        if out and rng.random() < repeat_p:
            s = rng.randrange(len(out))
            ln = rng.randrange(1, 40)
            out += out[s:s + ln]
        else:
            out.append(rng.choice(alphabet))
    return bytes(out[:n]) + b"$"


CASES = [
    dict(alphabet=b"ACGT", flags=dict(is_dna=True)),
    dict(alphabet=b"ACGTNacgtn%", flags=dict(is_dna=True)),
    dict(alphabet=b"ACGTNacgtn%", flags=dict(is_dna=True, allow_ambiguity=True)),
    dict(alphabet=b"ACGTNacgtn%", flags=dict(is_dna=True, ignore_softmask=True)),
    dict(alphabet=b"ACGTNacgtn%", flags=dict(is_dna=True, allow_ambiguity=True, ignore_softmask=True)),
    dict(alphabet=b"ACDEFGHIKLMNPQRSTVWY%", flags=dict()),
    dict(alphabet=b"AB", flags=dict()),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("seed", range(4))
def test_full_sort_matches_spec(case, seed):
    c = CASES[case]
    rng = random.Random(1000 * case + seed)
    text = rand_text(rng, rng.randrange(5, 700), c["alphabet"], repeat_p=0.05)
    res = O.oracle_build(text, num_partitions=rng.choice([1, 2, 5]), random_seed=seed + 1,
                         threads=rng.choice([1, 2]), **c["flags"])
    t, sa, lcp = O.spec_build(text, **c["flags"])
    assert res.text == t
    assert res.sa.tolist() == sa
    assert res.lcp.tolist() == lcp


@pytest.mark.parametrize("mask", ["101", "1101", "11011", "10111011", "1101101101", "111010010100110111"])
@pytest.mark.parametrize("case", [0, 2, 5, 6])
def test_seed_mask_matches_spec(mask, case):
    c = CASES[case]
    import zlib
    rng = random.Random(zlib.crc32(repr((mask, case)).encode()))
    text = rand_text(rng, rng.randrange(5, 600), c["alphabet"], repeat_p=0.05)
    res = O.oracle_build(text, num_partitions=rng.choice([1, 3]), random_seed=3, seed_mask=mask,
                         threads=2, **c["flags"])
    t, sa, lcp = O.spec_build(text, seed_mask_str=mask, **c["flags"])
    assert res.sa.tolist() == sa
    assert res.lcp.tolist() == lcp


@pytest.mark.parametrize("seed", range(4))
def test_mql_without_ties_equals_full_sort(seed):
    """SURVEY 8a rule 4: if no two indexed suffixes share Q bytes the MQL build equals the full sort."""
    rng = random.Random(seed)
    text = rand_text(rng, 500, b"ACDEFGHIKLMNPQRSTVWY")
    full = O.oracle_build(text, num_partitions=4)
    q = int(full.lcp.max()) + 1
    res = O.oracle_build(text, num_partitions=4, max_query_len=q)
    assert res.sa.tolist() == full.sa.tolist()
    assert res.lcp.tolist() == full.lcp.tolist()
    # one below the max LCP: order on the first Q bytes still exact, LCP values < Q exact
    q2 = max(1, q - 2)
    res2 = O.oracle_build(text, num_partitions=4, max_query_len=q2)
    t = res2.text
    keys = [t[p:p + q2] for p in res2.sa.tolist()]
    assert keys == sorted(keys)
    true = np.array([0] + [next((k for k in range(min(len(a), len(b))) if a[k] != b[k]), min(len(a), len(b)))
                           for a, b in zip([t[p:] for p in res2.sa.tolist()[:-1]], [t[p:] for p in res2.sa.tolist()[1:]])])
    small = res2.lcp < q2
    assert (res2.lcp[small] == true[small]).all()
    assert (true[~small] >= q2).all()
