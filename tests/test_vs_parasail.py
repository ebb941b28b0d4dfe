"""Opt-in differential tests against the REAL parasail (the reference's engine, pip parasail==1.3.4). parasail is not
installable in the build image (no network), so this module is skipped there; run it wherever parasail exists:
  * the oracle against parasail.ssw on random pairs, all eight scorings of the fuzz fixture, lower-case and N templates;
  * with a GPU (`-m gpu`): the CUDA path against parasail.ssw on the same pairs, through the C ABI;
  * where /root/reference is present too: the reference's own stage function re-run with the real parasail must
    reproduce tests/golden/stage_reference.json (which was made with the oracle standing in for parasail).
Until one of these has run, the arithmetic of parasail.ssw is PARITY UNPINNED."""
import json
import os
import random
import sys

import pytest

parasail = pytest.importorskip("parasail")

from oracle import oracle as O  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
SCORINGS = [(2, -2, 3, 1), (1, -2, 3, 1), (2, -3, 5, 2), (1, -1, 1, 1), (3, -2, 4, 1), (2, -2, 0, 1), (2, -2, 3, 0), (5, -4, 10, 1)]


def _cases(n=300, seed=5):
    rnd = random.Random(seed)
    for k in range(n):
        alpha = rnd.choice(["AC", "ACGT", "ACGTN", "AAAC", "ACGTacgt", "ACGTNRYn"])
        a = "".join(rnd.choice(alpha) for _ in range(rnd.randint(1, 500 if k % 10 else 3000)))
        b = list(a if rnd.random() < 0.5 else O.reverse_complement(a))
        for _ in range(rnd.randint(0, max(1, len(a) // 10))):
            if not b:
                break
            p = rnd.randrange(len(b))
            op = rnd.random()
            if op < 0.4:
                b[p] = rnd.choice(alpha)
            elif op < 0.7:
                del b[p]
            else:
                b.insert(p, rnd.choice(alpha))
        b = "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 800))) + "".join(b) + \
            "".join(rnd.choice(alpha) for _ in range(rnd.randint(0, 800)))
        yield a, b or "A"


def _ref_tuple(a, b, sc):
    m = parasail.matrix_create("ACGT", sc[0], sc[1])
    r = parasail.ssw(a, b, sc[2], sc[3], m)
    return None if r is None else (r.score1, r.read_begin1, r.read_end1, r.ref_begin1, r.ref_end1)


def test_oracle_equals_real_parasail():
    for k, sc in enumerate(SCORINGS):
        for a, b in _cases(120, seed=5 + k):
            ref = _ref_tuple(a, b, sc)
            got = O.ssw(a, b, sc[2], sc[3], sc[0], sc[1])
            if ref is None:
                assert got is None
                continue
            if ref[0] == 0:
                assert got.score1 == 0
                continue
            assert got.astuple() == ref, (sc, a, b)


def test_saturation_rule_equals_real_parasail():
    """None exactly when parasail returns None: scores around INT16_MAX - match - 1."""
    rnd = random.Random(3)
    for n in (16370, 16381, 16382, 16383, 16390):
        a = "".join(rnd.choice("ACGT") for _ in range(n))
        assert (O.ssw(a, a, engine="striped") is None) == (_ref_tuple(a, a, (2, -2, 3, 1)) is None), n


@pytest.mark.gpu
def test_cuda_path_equals_real_parasail(engine):
    from nanomonsv_b200 import ssw as gpu_ssw
    for k, sc in enumerate(SCORINGS):
        pairs = list(_cases(120, seed=5 + k))
        m = gpu_ssw.matrix_create("ACGT", sc[0], sc[1])
        got = gpu_ssw.ssw_batch(pairs, sc[2], sc[3], m, engine=engine)
        for (a, b), g in zip(pairs, got):
            ref = _ref_tuple(a, b, sc)
            if ref is None:
                assert g is None
            elif ref[0] == 0:
                assert g.score1 == 0
            else:
                assert (g.score1, g.read_begin1, g.read_end1, g.ref_begin1, g.ref_end1) == ref, (sc, a, b)


@pytest.mark.skipif(not os.path.isdir("/root/reference/nanomonsv"), reason="the reference checkout is not on this box")
def test_stage_fixture_is_reproduced_with_the_real_parasail(tmp_path):
    """The committed stage fixture was produced by the reference's stage function with the ORACLE standing in for
    parasail; with the real engine the reference must write the same files."""
    sys.path.insert(0, GOLDEN)
    try:
        import make_golden_stage as S
        fix = json.load(open(os.path.join(GOLDEN, "stage_reference.json")))
        ref = S.load_reference(real_parasail=parasail)
        S.REGISTRY["records"], S.REGISTRY["contigs"] = fix["records"], fix["contigs"]
        fsv, fsb = str(tmp_path / "refined_bp.txt"), str(tmp_path / "refined_bp.sbnd.txt")
        open(fsv, "w").write("".join(l + "\n" for l in fix["sv_file"]))
        open(fsb, "w").write("".join(l + "\n" for l in fix["sbnd_file"]))
        out = {k: str(tmp_path / k) for k in ("count", "info", "count_sbnd", "info_sbnd")}
        ref.count_sread_by_alignment(fsv, str(tmp_path / "tumor.bam"), out["count"], out["info"], str(tmp_path / "ref.fa"),
                                     sbnd_file=fsb, output_count_file_sbnd=out["count_sbnd"],
                                     output_alignment_info_file_sbnd=out["info_sbnd"], **fix["params"])
        import gc
        gc.collect()
        for k in out:
            assert sorted(open(out[k]).read().splitlines()) == sorted(fix["expected"][k]), k
    finally:
        sys.path.remove(GOLDEN)
        for mod in ("pysam", "nanomonsv", "nanomonsv.count_sread_by_alignment", "nanomonsv.my_seq", "nanomonsv.pyssw"):
            sys.modules.pop(mod, None)
        sys.modules["parasail"] = parasail


@pytest.mark.skipif(not os.path.isdir("/root/reference/nanomonsv"), reason="the reference checkout is not on this box")
def test_callsite_fixtures_are_reproduced_with_the_real_parasail():
    """locate_bp_sbnd, generate_paf_file and filter_sv_insertion_match of the reference, re-run with the real parasail on the
    generator's inputs, must write what tests/golden/callsites_reference.json holds (made with the oracle standing in)."""
    sys.path.insert(0, GOLDEN)
    try:
        import make_golden_callsites as M
        fix = json.load(open(os.path.join(GOLDEN, "callsites_reference.json")))
        got = M.build(real_parasail=parasail)
        assert got["sbnd"]["output"] == fix["sbnd"]["output"] and got["sbnd"]["log"] == fix["sbnd"]["log"]
        assert got["paf"]["paf"] == fix["paf"]["paf"] and got["paf"]["parasail_error"] == fix["paf"]["parasail_error"]
        assert got["filter"]["filters_after"] == fix["filter"]["filters_after"]
    finally:
        sys.path.remove(GOLDEN)
        for mod in [k for k in sys.modules if k == "pysam" or k == "nanomonsv" or k.startswith("nanomonsv.")]:
            sys.modules.pop(mod, None)
        sys.modules["parasail"] = parasail
