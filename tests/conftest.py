import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GOLDEN = ROOT / "tests" / "golden"

# (golden file, input fasta, create flags) -- flags from the reference's mk_test_files.py:63-87
# and Makefile:40-41 (1s.sufr).  Defaults -n 16 -r 42 -D % (sufr/src/lib.rs:91-124).
GOLDEN_CASES = [
    ("1.sufr", "1.fa", dict(is_dna=True)),
    ("2.sufr", "2.fa", dict(is_dna=True)),
    ("3.sufr", "3.fa", dict(is_dna=True)),
    ("2d.sufr", "2.fa", dict(is_dna=True, delimiter=b"N")),
    ("abba.sufr", "abba.fa", dict()),
    ("1n.sufr", "1.fa", dict(is_dna=True, allow_ambiguity=True)),
    ("2n.sufr", "2.fa", dict(is_dna=True, allow_ambiguity=True)),
    ("1s.sufr", "1.fa", dict(is_dna=True, ignore_softmask=True)),
    ("2s.sufr", "2.fa", dict(is_dna=True, ignore_softmask=True)),
    ("2ns.sufr", "2.fa", dict(is_dna=True, allow_ambiguity=True, ignore_softmask=True)),
    ("long_dna_sequence.sufr", "long_dna_sequence.fa", dict(is_dna=True)),
    ("long_dna_sequence_allow_ambiguity.sufr", "long_dna_sequence.fa", dict(is_dna=True, allow_ambiguity=True)),
    ("uniprot.sufr", "uniprot.fa", dict()),
    ("uniprot-masked.sufr", "uniprot.fa", dict(seed_mask="10111011")),
]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
