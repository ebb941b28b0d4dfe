"""Regenerates tests/golden/callsites_reference.json by running the REFERENCE'S OWN code of three other parasail.ssw call sites
 from /root/reference, unmodified, on small synthetic inputs:

  * nanomonsv/locate_bp_sbnd.py:43-66   `locate_bp_sbnd(consensus_file, output_file, reference_fasta, debug)` with its helper
    get_refined_bp_sbnd (:14-40): matrix 1/-2, reference window vs the whole consensus;
  * nanomonsv/generate_consensus.py:100-127   `Consensus_generator.generate_paf_file(query_fasta, target_fasta, output_file)`:
    matrix 2/-2, whole reads (incl. reads longer than 16 383 bases) as the first sequence;
  * nanomonsv/post_proc.py:84-134   `Sv_filterer.filter_sv_insertion_match` on the reference's own `Sv` objects, driven in the pair
    order of `apply_filters` (:226-232): junction sequence vs insertion sequence, both strands, matrix 2/-2.

`pysam` and `parasail` are absent in this image and are replaced by stand-ins registered in sys.modules before the import
(the same arrangement as make_golden_stage.py): pysam.FastaFile over in-memory contigs, parasail.ssw with numbers from this
repository's CPU oracle. The fixture therefore pins the reference's host logic around the call (windows, clipping, strand
handling, file parsing, position arithmetic, output and log formats) on the reference itself; the arithmetic inside
parasail.ssw remains the oracle's restatement. Inputs and produced outputs only are stored.

    python tests/golden/make_golden_callsites.py          (build container, /root/reference present)
"""
import importlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from nanomonsv_b200.my_seq import reverse_complement    # noqa: E402
from oracle import oracle as O                          # noqa: E402

CONTIGS = {}


class _FastaFile:
    def __init__(self, path):
        pass

    def fetch(self, chrom, start, end):
        return CONTIGS[chrom][max(start, 0):max(end, 0)]

    def get_reference_length(self, chrom):
        return len(CONTIGS[chrom])

    def close(self):
        pass


class _Matrix:
    def __init__(self, alphabet, match, mismatch):
        assert alphabet == "ACGT"
        self.match, self.mismatch = match, mismatch


def _ssw(s1, s2, gap_open, gap_extend, matrix):
    return O.ssw(s1, s2, gap_open, gap_extend, matrix.match, matrix.mismatch)


REAL_PARASAIL = None      # set by build(real_parasail=...): tests/test_vs_parasail.py re-runs the reference with the real engine


def load(module):
    pysam = types.ModuleType("pysam")
    pysam.FastaFile = _FastaFile
    pysam.AlignmentFile = object
    parasail = REAL_PARASAIL
    if parasail is None:
        parasail = types.ModuleType("parasail")
        parasail.matrix_create, parasail.ssw = (lambda a, m, x: _Matrix(a, m, x)), _ssw
    sys.modules["pysam"], sys.modules["parasail"] = pysam, parasail
    for name in [k for k in sys.modules if k == "nanomonsv" or k.startswith("nanomonsv.")]:
        del sys.modules[name]        # a module imported earlier holds the parasail it was imported with
    pkg = types.ModuleType("nanomonsv")
    pkg.__path__ = ["/root/reference/nanomonsv"]
    sys.modules["nanomonsv"] = pkg
    return importlib.import_module("nanomonsv." + module)


def rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))


def mutate(rng, s, rate):
    out = []
    for ch in s:
        r = rng.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append("ACGT"[rng.integers(0, 4)]); continue
        out.append(ch)
        if r > 1 - rate / 3:
            out.append("ACGT"[rng.integers(0, 4)])
    return "".join(out)


def sbnd_case(rng, tmp):
    """Consensus lines keyed chr,start,end,dir,_,id around breakpoints of two contigs, incl. windows clipped at both contig ends,
    a soft-masked reference stretch, a consensus unrelated to its window and one longer than 1000 bases."""
    CONTIGS.clear()
    CONTIGS["chrA"] = rand_seq(rng, 6000)
    c = rand_seq(rng, 3000)
    CONTIGS["chrB"] = c[:1200] + c[1200:1500].lower() + c[1500:]
    lines = []
    specs = [("chrA", 2000, 2030, '+'), ("chrA", 3100, 3140, '-'), ("chrA", 40, 90, '+'), ("chrA", 5950, 6100, '-'),
             ("chrB", 1300, 1340, '+'), ("chrB", 1400, 1420, '-'), ("chrA", 0, 30, '+'), ("chrB", 2990, 3050, '+'),
             ("chrA", 4000, 4040, '+'), ("chrA", 4500, 4530, '-')]
    for k, (ch, s, e, d) in enumerate(specs):
        ref = CONTIGS[ch].upper()
        bp = (s + e) // 2
        if d == '+':      # the contig continues from the reference bases left of the breakpoint into foreign sequence
            left = ref[max(bp - int(rng.integers(80, 400)), 0):bp]
            cons = mutate(rng, left, 0.02) + rand_seq(rng, int(rng.integers(100, 1300 if k == 8 else 500)))
        else:
            right = ref[bp - 1:bp - 1 + int(rng.integers(80, 400))]
            cons = mutate(rng, reverse_complement(right), 0.02) + rand_seq(rng, int(rng.integers(100, 500)))
        if k == 9:
            cons = rand_seq(rng, 300)
        lines.append(f"{ch},{s},{e},{d},sbnd,{k}\t{cons}")
    cfile = os.path.join(tmp, "consensus.txt")
    open(cfile, "w").write("\n".join(lines) + "\n")
    mod = load("locate_bp_sbnd")
    out = os.path.join(tmp, "out.txt")
    mod.locate_bp_sbnd(cfile, out, "unused.fa", True)
    return {"contigs": dict(CONTIGS), "consensus_lines": lines, "output": open(out).read().splitlines(),
            "log": open(out + ".locate_bp.sbnd.log").read().splitlines()}


def paf_case(rng, tmp):
    """Supporting reads of one candidate against its first consensus: reads that hold the consensus (either complete or running
    off one end), unrelated reads, reads longer than 16 383 bases (the 16-bit limit is about the SCORE, not the length), a
    header with blanks, lower-case bases, and a target file with two records (the last one counts, id cut at the blank)."""
    cons = rand_seq(rng, 1800)
    reads = []
    for k in range(14):
        n_left, n_right = int(rng.integers(0, 9000)), int(rng.integers(0, 9000))
        if k == 3:
            n_left, n_right = 12000, 9000         # 22 kb read
        if k == 4:
            n_left, n_right = 17000, 200
        body = mutate(rng, cons, 0.08)
        if k == 5:
            body = body[:900]                      # the read ends inside the consensus
        if k == 6:
            body = body[700:]
        if k in (7, 8):
            body = rand_seq(rng, 1500)             # unrelated
        seq = rand_seq(rng, n_left) + body + rand_seq(rng, n_right)
        if k == 9:
            seq = seq[:50] + seq[50:400].lower() + seq[400:]
        if k == 10:
            seq = seq[:300] + "N" * 25 + seq[325:]
        name = f"read{k} extra words" if k == 2 else f"read{k}"
        reads.append((name, seq))
    qf, tf, of = (os.path.join(tmp, n) for n in ("q.fa", "t.fa", "o.paf"))
    with open(qf, "w") as h:
        for name, seq in reads:
            h.write(f">{name}\n{seq}\n")
    with open(tf, "w") as h:
        h.write(f">first ignored\n{rand_seq(rng, 50)}\n>cons_1 second word\n{cons}\n")
    mod = load("generate_consensus")
    me = types.SimpleNamespace(parasail_error=[])
    n = mod.Consensus_generator.generate_paf_file(me, qf, tf, of)
    return {"query_fasta": open(qf).read(), "target_fasta": open(tf).read(), "paf": open(of).read().splitlines(),
            "returned": n, "parasail_error": [list(x) for x in me.parasail_error]}


def filter_case(rng):
    """SV list around two insertions: translocation-like junctions that are the two ends of the first insertion (one of them
    written from the other strand), a junction with an unrelated partner, one far away, a short-insert SV, an SV that arrives
    already filtered, and a second insertion a few bases away (its pairs are met after the first insertion has filtered them).
    The stage loop is the one of Sv_filterer.apply_filters (post_proc.py:226-232); every decision is the reference's own
    filter_sv_insertion_match (:84-134) on the reference's own Sv objects."""
    import itertools
    CONTIGS.clear()
    CONTIGS["chrA"] = rand_seq(rng, 8000)
    b = rand_seq(rng, 8000)
    CONTIGS["chrB"] = b[:5100] + b[5100:5200].lower() + b[5200:]
    src = CONTIGS["chrB"].upper()[5000:5300]
    ins1_seq = mutate(rng, src, 0.02)
    svs = [  # chr1, pos1, dir1, chr2, pos2, dir2, inseq, id, initial filter
        ("chrA", 3003, "+", "chrB", 5001, "-", "---"[:0], "t1", []),
        ("chrA", 2990, "+", "chrB", 7000, "-", "", "t2", []),
        ("chrA", 6000, "+", "chrB", 5001, "-", "", "t3", []),
        ("chrA", 3000, "+", "chrA", 3001, "-", ins1_seq, "ins1", []),
        ("chrB", 5300, "+", "chrA", 3001, "-", "", "t4", []),
        ("chrA", 3001, "-", "chrB", 5300, "+", "", "t5", []),
        ("chrA", 3010, "+", "chrA", 3011, "-", rand_seq(rng, 80), "ins2", []),
        ("chrA", 3005, "+", "chrB", 5003, "-", "ACGTACGTAC", "t6", []),
        ("chrA", 3002, "+", "chrB", 5001, "-", "", "t7", ["Too_low_VAF"]),
        ("chrA", 2998, "+", "chrB", 5000, "-", rand_seq(rng, 40), "t8", []),
    ]
    mod = load("post_proc")
    objs = []
    for c1, p1, d1, c2, p2, d2, inseq, sid, flt in svs:
        o = mod.Sv(c1, p1, d1, c2, p2, d2, inseq, sid, 30, 10, 30, 0)
        o.filter = list(flt)
        objs.append(o)
    f = object.__new__(mod.Sv_filterer)          # without __init__ (it opens files through pysam)
    f.reference_h, f.min_indel_size, f.bp_dist_margin, f.validate_seg_len = _FastaFile(None), 50, 30, 100
    for a, c in itertools.combinations(objs, 2):
        if len(a.filter) > 0 or len(c.filter) > 0:
            continue
        if len(a.inseq) < f.min_indel_size and len(c.inseq) >= f.min_indel_size:
            f.filter_sv_insertion_match(a, c)
        elif len(c.inseq) < f.min_indel_size and len(a.inseq) >= f.min_indel_size:
            f.filter_sv_insertion_match(c, a)
    f.hout = open(os.devnull, "w")                # __del__ closes these
    return {"contigs": dict(CONTIGS), "svs": [list(x[:8]) + [x[8]] for x in svs], "filters_after": [o.filter for o in objs]}


def build(real_parasail=None) -> dict:
    global REAL_PARASAIL
    REAL_PARASAIL = real_parasail
    rng = np.random.default_rng(20261020)
    with tempfile.TemporaryDirectory() as tmp:
        fix = {"generator": "tests/golden/make_golden_callsites.py", "reference": "nanomonsv v0.8.0 (/root/reference), modules "
               "locate_bp_sbnd and generate_consensus imported unmodified; pysam / parasail stand-ins (parasail numbers from the oracle)",
               "sbnd": sbnd_case(rng, tmp), "paf": paf_case(rng, tmp), "filter": filter_case(rng)}
    return fix


def main():
    fix = build()
    json.dump(fix, open(os.path.join(HERE, "callsites_reference.json"), "w"), indent=0)
    print("sbnd lines", len(fix["sbnd"]["output"]), "of", len(fix["sbnd"]["consensus_lines"]), "; paf lines", len(fix["paf"]["paf"]),
          "returned", fix["paf"]["returned"], "; filters", fix["filter"]["filters_after"])


if __name__ == "__main__":
    main()
