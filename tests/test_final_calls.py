"""Final calls: the reference's own integrate_realignment_result (nanomonsv/post_proc.py:264-303) on
the count files of the validation stage. tests/golden/integrate_reference.json holds the count files of the small
validation fixture and the result.txt the reference made from them (generator: tests/golden/make_golden_integrate.py).
  * CPU: the fixture is consistent with validate_small.json and, where /root/reference is present, reproducible by
    re-running the reference function;
  * GPU: the count files WRITTEN BY THE CUDA STAGE FUNCTION are byte-identical to the ones the reference function was
    fed, so its result.txt is the final-call list of a `nanomonsv get` whose validation ran on the GPU."""
import json
import os
import sys

import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _fixtures():
    return (json.load(open(os.path.join(GOLDEN, "integrate_reference.json"))),
            json.load(open(os.path.join(GOLDEN, "validate_small.json"))))


def test_final_call_fixture_is_made_from_the_stage_fixture():
    integ, small = _fixtures()
    assert integ["tumor_count"] == small["samples"]["tumor"]["count"]
    assert integ["control_count"] == small["samples"]["control"]["count"]
    by_name = {c["name"]: c for c in integ["cases"]}
    res = by_name["tumor_and_control"]["result"]
    assert res[0].split("\t")[-5:] == ["Checked_Read_Num_Tumor", "Supporting_Read_Num_Tumor", "Checked_Read_Num_Control",
                                       "Supporting_Read_Num_Control", "Is_Filter"]
    # the read-count rules of post_proc.py:289-295 on these files: >= 3 supporting tumor reads, <= 1 control read
    keep = {tuple(l.split("\t")[:8]) for l in integ["tumor_count"] if int(l.split("\t")[9]) >= 3}
    assert {tuple(l.split("\t")[:8]) for l in res[1:]} <= keep and len(res) - 1 == 4
    assert len(by_name["tumor_and_control_min2"]["result"]) - 1 == 7


@pytest.mark.skipif(not os.path.isdir("/root/reference/nanomonsv"), reason="the reference checkout is not on this box")
def test_reference_function_reproduces_the_committed_final_calls():
    sys.path.insert(0, GOLDEN)
    try:
        import make_golden_integrate as M
        integ, small = _fixtures()
        for case in integ["cases"]:
            got = M.run_reference(small["contigs"], integ["tumor_count"], integ["control_count"] if case["control"] else None,
                                  **case["params"])
            assert got == case["result"], case["name"]
    finally:
        sys.path.remove(GOLDEN)
        for mod in ("pysam", "parasail", "nanomonsv", "nanomonsv.post_proc", "nanomonsv.count_sread_by_alignment",
                    "nanomonsv.my_seq", "nanomonsv.logger", "nanomonsv.pyssw"):
            sys.modules.pop(mod, None)


@pytest.mark.gpu
def test_gpu_count_files_are_the_inputs_of_the_reference_final_calls(engine, tmp_path):
    from nanomonsv_b200 import count_sread_by_alignment as C
    from oracle import oracle as O
    integ, small = _fixtures()
    fasta = O.DictFasta(small["contigs"])
    written = {}
    for sample, d in small["samples"].items():
        seam = tmp_path / (sample + ".tsv")
        seam.write_text("\n".join(d["seam"]) + "\n")
        C.count_sread_from_seam(str(seam), str(seam) + ".count", str(seam) + ".info", fasta, False, 40, 1.2, engine=engine)
        written[sample] = open(str(seam) + ".count").read().splitlines()
    assert written["tumor"] == integ["tumor_count"] and written["control"] == integ["control_count"]
    calls = [l for l in {c["name"]: c for c in integ["cases"]}["tumor_and_control"]["result"][1:] if l.endswith("PASS")]
    assert len(calls) == 3 and len({c["name"]: c for c in integ["cases"]}["tumor_and_control"]["result"]) - 1 == 4
