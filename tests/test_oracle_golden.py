"""Pins the CPU oracle (oracle/sufr_oracle.cpp) against the reference's own golden `.sufr`
files and in-source known-answer vectors."""
import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_CASES
import oracle as O
from sufrfile import parse_sufr


@pytest.mark.parametrize("golden,fasta,flags", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
@pytest.mark.parametrize("threads", [1, 3])
def test_golden_whole_file(golden, fasta, flags, threads):
    flags = dict(flags)
    delim = flags.pop("delimiter", b"%")
    seq = O.read_sequence_file(GOLDEN / "inputs" / fasta, delim)
    res = O.oracle_build(seq.seq, sequence_starts=seq.start_positions, sequence_names=seq.sequence_names,
                         num_partitions=16, random_seed=42, threads=threads, **flags)
    want = (GOLDEN / "expected" / golden).read_bytes()
    g = parse_sufr(want)
    assert res.text == g.text
    np.testing.assert_array_equal(res.sa, g.sa)
    np.testing.assert_array_equal(res.lcp, g.lcp)
    assert res.file_bytes == want


@pytest.mark.parametrize("golden,fasta,flags", GOLDEN_CASES, ids=[c[0] for c in GOLDEN_CASES])
def test_golden_invariant_under_partitions_and_seed(golden, fasta, flags):
    """Full-sort and mask outputs do not depend on pivots / partition count."""
    flags = dict(flags)
    delim = flags.pop("delimiter", b"%")
    seq = O.read_sequence_file(GOLDEN / "inputs" / fasta, delim)
    g = parse_sufr((GOLDEN / "expected" / golden).read_bytes())
    for n, r in ((1, 1), (3, 7), (64, 12345)):
        res = O.oracle_build(seq.seq, sequence_starts=seq.start_positions, sequence_names=seq.sequence_names,
                             num_partitions=n, random_seed=r, threads=2, **flags)
        np.testing.assert_array_equal(res.sa, g.sa)
        np.testing.assert_array_equal(res.lcp, g.lcp)


def test_read_sequence_file_kat():
    # libsufr/src/util.rs:183-194
    d = O.read_sequence_file(GOLDEN / "inputs" / "2.fa", b"N")
    assert d.seq == b"ACGTacgtNacgtACGT$"
    assert d.start_positions == [0, 9]
    assert d.sequence_names == ["ABC", "DEF"]


def test_empty_input_is_error():
    # sufr/tests/cli.rs:102-109 (create on an empty/missing input fails)
    with pytest.raises(O.OracleError):
        O.read_sequence_file(GOLDEN / "inputs" / "empty.fa")


def test_lib_rs_suffix_file_32():
    # libsufr/src/lib.rs:45-92
    d = O.read_sequence_file(GOLDEN / "inputs" / "2.fa", b"N")
    r = O.oracle_build(d.seq, is_dna=True, num_partitions=2, random_seed=0, index_bits=32,
                       sequence_starts=d.start_positions, sequence_names=d.sequence_names)
    assert r.text == b"ACGTACGTNACGTACGT$"
    assert r.sa.tolist() == [17, 13, 9, 0, 4, 14, 10, 1, 5, 15, 11, 2, 6, 16, 12, 3, 7]
    assert r.lcp.tolist() == [0, 0, 4, 8, 4, 0, 3, 7, 3, 0, 2, 6, 2, 0, 1, 5, 1]
    assert r.sa.dtype == np.uint32


def test_lib_rs_suffix_file_64():
    # libsufr/src/lib.rs:94-140 (u64 index width, allow_ambiguity)
    d = O.read_sequence_file(GOLDEN / "inputs" / "1.fa", b"N")
    r = O.oracle_build(d.seq, is_dna=True, allow_ambiguity=True, num_partitions=2, random_seed=0,
                       index_bits=64, sequence_starts=d.start_positions, sequence_names=d.sequence_names)
    assert r.text == b"ACGTNNACGT$"
    assert r.num_suffixes == 11
    assert r.sa.tolist() == [10, 6, 0, 7, 1, 8, 2, 5, 4, 9, 3]
    assert r.lcp.tolist() == [0, 0, 4, 0, 3, 0, 2, 0, 1, 0, 1]
    assert r.sa.dtype == np.uint64
    g = parse_sufr(r.file_bytes) if False else None  # u64 files with text_len < 2^32 are not reader-parsable


def test_lib_rs_smol_num_suffixes():
    # libsufr/src/lib.rs:142-173
    d = O.read_sequence_file(GOLDEN / "inputs" / "smol.fa", b"N")
    r = O.oracle_build(d.seq, is_dna=True, num_partitions=2, random_seed=0,
                       sequence_starts=d.start_positions, sequence_names=d.sequence_names)
    assert r.num_suffixes == 364
    # subsample_suffix_array(mql) keeps lcp < mql (sufr_file.rs:443-453); lib.rs:175-216
    for q, count in ((1, 5), (2, 20), (3, 71), (5, 293)):
        keep = r.lcp < q
        assert int(keep.sum()) == count
    assert r.sa[r.lcp < 1].tolist() == [365, 364, 92, 224, 363]
    assert np.nonzero(r.lcp < 1)[0].tolist() == [0, 1, 94, 191, 284]


@pytest.mark.parametrize("fasta,mask,num,want", [
    # libsufr/src/lib.rs:221-264, 266-321, 324-365
    ("mostlya1.fa", "101", 8, [7, 6, 5, 4, 2, 0, 1, 3]),
    ("mostlya2.fa", "11011", 17, [16, 13, 9, 5, 1, 12, 8, 4, 0, 14, 10, 6, 2, 15, 11, 7, 3]),
    ("spaced_input.fa", "11000111", 43,
     [42, 18, 12, 0, 32, 29, 13, 23, 21, 6, 40, 1, 33, 19, 30, 10, 28, 9, 17, 14, 4, 26, 39, 22, 25, 38, 24,
      35, 7, 36, 15, 41, 5, 20, 31, 11, 27, 8, 16, 3, 37, 34, 2]),
])
def test_lib_rs_spaced_seeds(fasta, mask, num, want):
    d = O.read_sequence_file(GOLDEN / "inputs" / fasta, b"N")
    r = O.oracle_build(d.seq, is_dna=True, num_partitions=1, random_seed=0, seed_mask=mask,
                       sequence_starts=d.start_positions, sequence_names=d.sequence_names)
    assert r.num_suffixes == num
    assert r.sa.tolist() == want


def _mk(text, **kw):
    base = dict(is_dna=False, num_partitions=2, random_seed=0, index_bits=32)
    base.update(kw)
    return O.Oracle(text, **base)


def test_builder_is_less():
    # sufr_builder.rs:1041-1081
    o = _mk(b"TTTAGC")
    assert o.is_less(1, 0) and not o.is_less(0, 1) and not o.is_less(2, 3) and o.is_less(3, 0)


def test_builder_is_less_max_query_len():
    # sufr_builder.rs:1083-1126
    o = _mk(b"TTTAGC", max_query_len=2)
    assert not o.is_less(1, 0) and not o.is_less(0, 1) and not o.is_less(2, 3) and o.is_less(3, 0)


def test_builder_is_less_seed_mask():
    # sufr_builder.rs:1128-1172
    o = _mk(b"TTTTAT", seed_mask="101")
    assert not o.is_less(0, 1) and not o.is_less(1, 0) and not o.is_less(0, 3) and not o.is_less(3, 0)


def test_builder_find_lcp_no_seed_mask():
    # sufr_builder.rs:1174-1220
    o = _mk(b"TTTAGC")
    assert o.find_lcp(0, 1, 6, 0) == 2
    assert o.find_lcp(0, 2, 6, 0) == 1
    assert o.find_lcp(0, 1, 1, 0) == 1
    assert o.find_lcp(0, 3, 6, 0) == 0


def test_builder_find_lcp_with_seed_mask():
    # sufr_builder.rs:1222-1258
    o = _mk(b"TTTTTA", seed_mask="1101", random_seed=42)
    assert o.find_lcp(0, 1, 3, 0) == 3
    assert o.find_lcp(0, 2, 3, 0) == 2
    assert o.find_lcp(0, 5, 3, 0) == 0


def test_builder_upper_bound_1():
    # sufr_builder.rs:1260-1293
    o = _mk(b"TTTAGC", random_seed=42)
    assert o.upper_bound(3, [5, 4]) == 0
    assert o.upper_bound(2, [3, 4, 5]) == 3
    assert o.upper_bound(5, [3, 4, 5]) == 1


def test_builder_upper_bound_2():
    # sufr_builder.rs:1295-1348 (u64 builder)
    o = _mk(b"ACGTNNACGT", random_seed=42, index_bits=64)
    assert o.upper_bound(0, [0]) == 0
    assert o.upper_bound(0, [6]) == 1
    assert o.upper_bound(6, [0]) == 0
    assert o.upper_bound(6, [6]) == 0
    assert o.upper_bound(0, [7, 8]) == 0
    assert o.upper_bound(1, [7, 8]) == 1
    assert o.upper_bound(9, [7, 8]) == 2
    assert o.upper_bound(9, [3]) == 0


def test_builder_upper_bound_seed_mask():
    # sufr_builder.rs:1350-1405
    o = _mk(b"ACGTNNACGT", random_seed=42, seed_mask="101")
    assert o.upper_bound(0, [0]) == 0
    assert o.upper_bound(0, [6]) == 0
    assert o.upper_bound(6, [0]) == 0
    assert o.upper_bound(6, [6]) == 0
    assert o.upper_bound(0, [7, 8]) == 0
    assert o.upper_bound(1, [7, 8]) == 0
    assert o.upper_bound(8, [7, 8]) == 1
    assert o.upper_bound(9, [7, 8]) == 2
    assert o.upper_bound(9, [3]) == 0


def test_find_lcp_full_offset():
    # util.rs:289-314
    assert [O.find_lcp_full_offset(i, "101") for i in range(3)] == [0, 2, 3]
    assert [O.find_lcp_full_offset(i, "11011") for i in range(5)] == [0, 1, 3, 4, 5]
    assert [O.find_lcp_full_offset(i, "10011001") for i in range(5)] == [0, 3, 4, 7, 8]


def test_seed_mask_types():
    # types.rs:62-78, 634-737
    for bad in ["", "0", "01", "10", "11", "0101", "1010", "1021", "1111", "abc", "1", "111", "00",
                "0111", "11100", "1a01"]:
        assert not O.seed_mask_valid(bad), bad
        assert O.seed_mask(bad) is None
    for good in ["101", "1001", "1101", "10101", "1110110110100001"]:
        assert O.seed_mask_valid(good), good
    b, p, d, w = O.seed_mask("110110101")
    assert (b, p, d, w) == ([1, 1, 0, 1, 1, 0, 1, 0, 1], [0, 1, 3, 4, 6, 8], [0, 0, 1, 1, 2, 3], 6)
    b, p, d, w = O.seed_mask("11101101101000011")
    assert p == [0, 1, 2, 4, 5, 7, 8, 10, 15, 16] and d == [0, 0, 0, 1, 1, 2, 2, 3, 7, 7] and w == 10


def test_argument_errors():
    # sufr_builder.rs:163-165, types.rs:81-83
    with pytest.raises(O.OracleError, match="Cannot use max_query_len and seed_mask together"):
        O.oracle_build(b"ACGT$", max_query_len=3, seed_mask="101")
    with pytest.raises(O.OracleError, match="Invalid seed mask '111'"):
        O.oracle_build(b"ACGT$", seed_mask="111")


def test_header_layout_1_sufr():
    # suffix_array.rs:351-366 doc-test: 1.sufr is 172 bytes, version 6, 9 suffixes
    g = parse_sufr((GOLDEN / "expected" / "1.sufr").read_bytes())
    assert (g.version, g.num_suffixes, g.text_pos, g.sa_pos, g.lcp_pos) == (6, 9, 72, 83, 119)


def test_too_few_pivot_positions_is_reported_not_hung():
    """The reference loops forever when a DNA text has fewer ACGT$ positions than pivots
    (sufr_builder.rs:787-796); the oracle raises instead."""
    with pytest.raises(O.OracleError, match="loop forever"):
        O.oracle_build(b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNACG$", is_dna=True, num_partitions=16)
