"""Opt-in known-answer test: the reference's disabled `test_validate`
(bk/test_main.py.bk:504-521) expects 46 checked / 10 supporting tumor reads and 37 / 0 control reads for sv_3 of its
test BAMs. The BAMs, GRCh38 and pysam are not in this image (.MISSING_LARGE_BLOBS), so the test runs only where

    NMSV_TEST_TUMOR_BAM, NMSV_TEST_CTRL_BAM, NMSV_TEST_GRCH38

point to tests/resource/bam/test_tumor.bam, test_ctrl.bam and the GRCh38 FASTA of the reference's test setup, pysam is
importable and a GPU is present. It drives the drop-in stage function exactly like run.py:501-511 (validate_main):
var_read_min_mapq = 40 (arg_parser.py:185), score ratio 1.2."""
import os

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
NEED = ("NMSV_TEST_TUMOR_BAM", "NMSV_TEST_CTRL_BAM", "NMSV_TEST_GRCH38")


def test_validate_known_answer(engine, tmp_path):
    if not all(os.environ.get(k) and os.path.exists(os.environ[k]) for k in NEED):
        pytest.skip("the reference's test BAMs / GRCh38 are not on this box (set " + ", ".join(NEED) + ")")
    pytest.importorskip("pysam")
    from nanomonsv_b200.count_sread_by_alignment import count_sread_by_alignment
    lines = [l for l in open(os.path.join(HERE, "golden", "validate_known_answer.tsv")).read().splitlines() if not l.startswith("#")]
    header, rec = lines[0].split("\t"), lines[1].split("\t")
    sv_list = tmp_path / "sv_list.txt"
    sv_list.write_text("\t".join(header[:8]) + "\n" + "\t".join(rec[:8]) + "\n")
    got = {}
    for sample, bam in (("tumor", os.environ[NEED[0]]), ("control", os.environ[NEED[1]])):
        fc, fi = str(tmp_path / f"{sample}.sread_count.txt"), str(tmp_path / f"{sample}.sread_info.txt")
        count_sread_by_alignment(str(sv_list), bam, fc, fi, os.environ[NEED[2]], var_read_min_mapq=40, use_ssw_lib=False,
                                 sort_option="-S 2G", debug=False, engine=engine)      # arg_parser.py:203
        f = open(fc).read().splitlines()[0].split("\t")
        assert f[:8] == rec[:8]
        got[sample] = (int(f[8]), int(f[9]))
    assert got["tumor"] == (int(rec[8]), int(rec[9])) == (46, 10)
    assert got["control"] == (int(rec[10]), int(rec[11])) == (37, 0)
