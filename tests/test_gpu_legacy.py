"""GPU parity test of the legacy-module wrapper (reference nanomonsv/long_read_validate.py): same alignments as the
live path, different host rules (key by insertion length, no upper-casing, <100 short del/dup rule, the combined
predicate that tests ref1 twice). The expectation is restated here from long_read_validate.py:156-417 with the CPU
oracle as the aligner."""
import dataclasses

import pytest

from nanomonsv_b200 import long_read_validate as LRV
from nanomonsv_b200 import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


class LowerCaseFasta:
    """Soft-masks a stretch of every contig (lower case), like a repeat-masked reference genome."""

    def __init__(self, inner, period=211, run=23):
        self.inner, self.period, self.run = inner, period, run

    def fetch(self, chrom, start, end):
        s = self.inner.fetch(chrom, start, end)
        start = max(start, 0)
        return "".join(c.lower() if ((start + i) % self.period) < self.run else c for i, c in enumerate(s))


def _expected(sv_rows, key2reads, key2mapq, fasta, min_mapq, L, ratio, st, en, margin):
    cnt, cnt_all, info = {}, {}, {}
    for F in sv_rows:
        key = LRV.legacy_key(F)
        reads = key2reads.get(key)
        if not reads:
            continue
        t = LRV.legacy_templates(F, fasta, L)
        dedup = {}
        for rid, seq in reads:
            dedup[rid] = seq
        rl = list(dedup.items())
        v1, v2 = O.ssw_check(t[0], rl, "striped"), O.ssw_check(t[1], rl, "striped")
        short = LRV.is_short_del_dup(key)
        if short:
            r1, r2 = O.ssw_check(t[2], rl, "striped"), O.ssw_check(t[3], rl, "striped")
        names = list(v1.keys())

        def cover(a, ln):
            return a[0] > ratio * ln and a[1] < st * ln and a[2] > en * ln

        if short:
            supp = [n for n in names if
                    (cover(v1[n], len(t[0])) and v1[n][0] >= r1[n][0] + margin and v1[n][0] >= r2[n][0] + margin) or
                    (cover(v2[n], len(t[1])) and v2[n][0] >= r1[n][0] + margin and v2[n][0] >= r1[n][0] + margin)]
        else:
            supp = [n for n in names if cover(v1[n], len(t[0])) or cover(v2[n], len(t[1]))]
        mq = key2mapq[key]
        supp = [n for n in supp if n in mq and mq[n][0] is not None and mq[n][0] >= min_mapq
                and mq[n][1] is not None and mq[n][1] >= min_mapq]
        cnt[key], cnt_all[key] = len(supp), len(names)
        info[key] = {n: v1[n] + v2[n] for n in supp}
    return cnt, cnt_all, info


def test_legacy_wrapper_matches_restated_reference_rules(engine, tmp_path):
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=12, coverage=7.0, read_len_mean=2000,
                              read_len_min=700, read_len_max=5000, contig_len=120000, n_contigs=2, seed=91,
                              short_fraction=0.7, sv_mix=(("DEL", 0.6), ("INS", 0.2), ("DUP", 0.2)))
    wl = synth.make_workload(cfg, samples=("tumor",))
    fasta = LowerCaseFasta(wl.fasta)
    sv_rows, key2reads, key2mapq = [], {}, {}
    for sv, reads in zip(wl.svs, wl.samples["tumor"]):
        F = sv.key.split(",")
        sv_rows.append(F)
        key = LRV.legacy_key(F)
        key2reads[key] = [(r.rid, synth.codes_to_str(r.codes)) for r in reads]
        key2mapq[key] = {r.rid: [r.mapq1, r.mapq2] for r in reads}
    assert any(LRV.is_short_del_dup(LRV.legacy_key(F)) for F in sv_rows)
    assert all(any(c.islower() for c in LRV.legacy_templates(F, fasta)[2]) for F in sv_rows)
    for params in ((0, 200, 1.4, 0.2, 0.8, 10), (40, 200, 1.2, 0.1, 0.9, 20)):
        min_mapq, L, ratio, st, en, margin = params
        got = LRV.long_read_validate_from_reads(sv_rows, key2reads, key2mapq, fasta, min_mapq, L, ratio, st, en, margin,
                                                engine=engine)
        exp = _expected(sv_rows, key2reads, key2mapq, fasta, min_mapq, L, ratio, st, en, margin)
        assert got[0] == exp[0] and got[1] == exp[1]
        assert got[2] == exp[2]
        assert sum(exp[0].values()) > 3


def test_legacy_ssw_check_switch(engine, tmp_path):
    q, t = tmp_path / "q.fa", tmp_path / "t.fa"
    q.write_text(">k\nACGTacgtTTGACCA\n")
    t.write_text(">r1\nGGACGTACGTTTGACCATT\n>r2\nTGGTCAAACGTACGTCC\n")
    exp = O.ssw_check("ACGTacgtTTGACCA", list(O.fasta_records(str(t))))
    assert LRV.ssw_check(str(q), str(t), False, engine=engine) == exp
    assert LRV.ssw_check(str(q), str(t), True, engine=engine) == exp
    assert LRV.ssw_check_parasail(str(q), str(t), engine=engine) == exp
