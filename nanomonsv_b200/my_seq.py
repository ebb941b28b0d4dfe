"""Host-side sequence helpers that the validation path needs (reference: nanomonsv/my_seq.py:20-26)."""
from __future__ import annotations

import numpy as np

# IUPAC complements for UPPER-CASE symbols only. Like the reference, every other character -- lower case
# included -- is carried over unchanged (only its position is reversed). That asymmetry matters downstream:
# the aligner's alphabet mapper folds case, so a lower-case 'a' in a template stays an A on the reverse strand.
_PAIRS = "AT CG GC TA WW SS MK KM RY YR BV VB DH HD NN".split()
_TABLE = {p[0]: p[1] for p in _PAIRS}
_TRANS = str.maketrans(_TABLE)


def reverse_complement(seq: str) -> str:
    return seq[::-1].translate(_TRANS)


# byte -> code: parasail.matrix_create("ACGT", ...) maps ACGT case-insensitively, everything else to a wildcard
CODE_LUT = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    CODE_LUT[ord(_c)] = _i
    CODE_LUT[ord(_c.lower())] = _i


CODE_TABLE = bytes(CODE_LUT.tolist())      # the same mapper for bytes.translate (3-4x faster than a numpy gather)


def encode(seq: str) -> np.ndarray:
    """ASCII sequence -> uint8 codes (0..3 = ACGT, 4 = wildcard); the array is a read-only view of a bytes object."""
    return np.frombuffer(seq.encode("latin-1").translate(CODE_TABLE), dtype=np.uint8)


def needs_explicit_revcomp(seq: str) -> bool:
    """True when complementing the *codes* would differ from encoding reverse_complement(seq): exactly when the
    sequence holds lower-case a/c/g/t, which reverse_complement leaves uncomplemented."""
    return any(c in seq for c in "acgt")
