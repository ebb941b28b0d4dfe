"""nanomonsv_b200 -- B200 (sm_100a) engine for nanomonsv's supporting-read validation hot path.

Host side mirrors the reference module nanomonsv/count_sread_by_alignment.py; the alignment, strand choice and
supporting-read decision run in hand-written CUDA behind the C ABI in include/nmsv.h. There is no CPU fallback.
"""
__version__ = "0.1.0"

from .engine import DEFAULT_SCORING, Engine, EngineError, Plan, SeqPack, default_engine  # noqa: F401
from .my_seq import reverse_complement  # noqa: F401
