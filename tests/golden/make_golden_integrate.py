"""Regenerates tests/golden/integrate_reference.json: the FINAL CALLS of the reference on the count files of the
validation stage. The reference's own nanomonsv/post_proc.py:264-303 (`integrate_realignment_result`: read-count and VAF
thresholds, then Sv_filterer -- duplicate / simple-repeat / insertion-match filters) is imported unmodified from
/root/reference with the stand-in pysam / parasail modules of make_golden_stage.py and run on the tumor and control
sread_count files of tests/golden/validate_small.json, with and without a control sample.

The CUDA stage function writes exactly these count files (tests/test_gpu_validate.py::test_golden_fixture holds them
byte for byte), so `result.txt` below is what `nanomonsv get` ends with when the validation runs on the GPU.
Nothing of the reference's source is written into this repository: inputs and produced outputs only.

    python tests/golden/make_golden_integrate.py          (build container, /root/reference present)
"""
import importlib
import json
import os
import tempfile

import make_golden_stage as S

HERE = os.path.dirname(os.path.abspath(__file__))


def run_reference(contigs, tumor_count, control_count, **kw):
    S.load_reference()                       # registers the stand-in pysam / parasail and the bare package
    S.REGISTRY["contigs"] = contigs
    post_proc = importlib.import_module("nanomonsv.post_proc")
    with tempfile.TemporaryDirectory() as tmp:
        ft, fc, out = os.path.join(tmp, "t.count"), os.path.join(tmp, "c.count"), os.path.join(tmp, "result.txt")
        open(ft, "w").write("".join(l + "\n" for l in tumor_count))
        if control_count is not None:
            open(fc, "w").write("".join(l + "\n" for l in control_count))
        post_proc.integrate_realignment_result(ft, fc if control_count is not None else None, out, os.path.join(tmp, "ref.fa"), **kw)
        import gc
        gc.collect()                         # Sv_filterer.__del__ closes the output file
        return open(out).read().splitlines()


def main():
    fix = json.load(open(os.path.join(HERE, "validate_small.json")))
    t, c = fix["samples"]["tumor"]["count"], fix["samples"]["control"]["count"]
    cases = []
    for name, ctrl, kw in (("tumor_and_control", c, {}), ("tumor_only", None, {}),
                           ("tumor_and_control_min2", c, dict(min_tumor_variant_read_num=2)),
                           ("tumor_only_min2_vaf20", None, dict(min_tumor_variant_read_num=2, min_tumor_VAF=0.2))):
        cases.append({"name": name, "control": ctrl is not None, "params": kw, "result": run_reference(fix["contigs"], t, ctrl, **kw)})
        print(name, len(cases[-1]["result"]) - 1, "calls")
    json.dump({"source": "nanomonsv/post_proc.py:264-303 run from /root/reference with stand-in pysam / parasail on the count "
                         "files of tests/golden/validate_small.json (tests/golden/make_golden_integrate.py)",
               "tumor_count": t, "control_count": c, "cases": cases},
              open(os.path.join(HERE, "integrate_reference.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
