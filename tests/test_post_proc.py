"""proc_sread_info_file (nanomonsv/post_proc.py:306-340) against outputs of the reference's own function
(tests/golden/proc_sread_info_vectors.json, generator tests/golden/make_golden_postproc.py)."""
import json
import os

import pytest

from nanomonsv_b200 import post_proc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
HEADER = "Chr_1\tPos_1\tDir_1\tChr_2\tPos_2\tDir_2\tInserted_Seq\tSV_ID\tChecked_Read_Num_Tumor\n"


def test_matches_the_reference_function(tmp_path):
    vec = json.load(open(os.path.join(GOLDEN, "proc_sread_info_vectors.json")))
    n = 0
    for case in vec["cases"]:
        fi, fr, fo = (str(tmp_path / x) for x in ("info.txt", "result.txt", "out.txt"))
        open(fi, "w").write("".join(l + "\n" for l in case["info"]))
        open(fr, "w").write(HEADER + "".join(k + "\t12\t5\n" for k in case["result_keys"]))
        post_proc.proc_sread_info_file(fi, fr, fo, case["validate_sequence_length"])
        assert open(fo).read().splitlines() == case["expected"], case["name"]
        n += len(case["expected"])
    assert n > 100


def test_quirks():
    f = "chr1\t100\t+\tchr2\t900\t-\t---\tid\tread\t700\t11\t390\t1000\t1385\t+\t700\t5\t395\t2000\t2400\t-".split("\t")
    assert post_proc.breakpoint_on_read(f, 200) == (int((385.0 / 379.0) * 189 + 1000), "+")        # tie: first segment
    f[9] = "699"
    assert post_proc.breakpoint_on_read(f, 200) == (int((-400.0 / 390.0) * 195 + 2400 + 3 + 1), "-")  # len('---') + 1
    f[16] = f[17] = "7"
    with pytest.raises(ZeroDivisionError):
        post_proc.breakpoint_on_read(f, 200)
