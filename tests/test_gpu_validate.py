"""GPU parity tests of the stage seam: per-SV checked / supporting counts and supporting-read info lines of the CUDA
path vs the CPU oracle's restatement of Alignment_counter, on the committed golden fixture and on seeded synthetic
workloads shaped like the BASELINE configs (scaled down so the oracle finishes in seconds)."""
import dataclasses
import json
import os

import pytest

from nanomonsv_b200 import count_sread_by_alignment as C
from nanomonsv_b200 import synth, _lib
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _run_both(tmp_path, wl, sample, engine, is_sbnd=False, L=200, min_mapq=40, ratio=1.2):
    seam = str(tmp_path / f"{sample}.seam.tsv")
    wl.write_seam(sample, seam)
    O.count_from_seam(seam, wl.fasta, seam + ".o_count", seam + ".o_info", is_sbnd=is_sbnd, var_read_min_mapq=min_mapq,
                      score_ratio_thres=ratio, validate_sequence_length=L, engine="striped")
    counter = C.count_sread_from_seam(seam, seam + ".g_count", seam + ".g_info", wl.fasta, is_sbnd, min_mapq, ratio,
                                      engine=engine, validate_sequence_length=L)
    exp_count = open(seam + ".o_count").read().splitlines()
    got_count = open(seam + ".g_count").read().splitlines()
    assert got_count == exp_count
    exp_info = sorted(open(seam + ".o_info").read().splitlines())
    got_info = sorted(open(seam + ".g_info").read().splitlines())
    assert got_info == exp_info
    return exp_count, exp_info, counter


def test_golden_fixture(engine, tmp_path):
    fix = json.load(open(os.path.join(GOLDEN, "validate_small.json")))
    fasta = O.DictFasta(fix["contigs"])
    n_supp = 0
    for sample, d in fix["samples"].items():
        seam = tmp_path / (sample + ".tsv")
        seam.write_text("\n".join(d["seam"]) + "\n")
        C.count_sread_from_seam(str(seam), str(seam) + ".count", str(seam) + ".info", fasta, False, 40, 1.2, engine=engine)
        assert open(str(seam) + ".count").read().splitlines() == d["count"]
        assert sorted(open(str(seam) + ".info").read().splitlines()) == d["info"]
        n_supp += len(d["info"])
    assert n_supp > 5


def test_canonical_with_short_del_dup_branch(engine, tmp_path):
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=14, coverage=9.0, read_len_mean=2500,
                              read_len_min=700, read_len_max=6000, contig_len=150000, n_contigs=3, seed=77,
                              short_fraction=0.6)
    wl = synth.make_workload(cfg)
    assert any(C.canonical_templates(s.key, wl.fasta, 200)[4] for s in wl.svs)
    total_supp = 0
    for sample in ("tumor", "control"):
        counts, info, _ = _run_both(tmp_path, wl, sample, engine)
        total_supp += len(info)
        if sample == "control":
            assert all(line.split("\t")[-1] in ("0", "1") for line in counts)
    assert total_supp >= 10


def test_long_templates_multi_tile(engine, tmp_path):
    """validate_sequence_length = 600 -> 1200-base templates: three 13-row-strip tiles / class selection."""
    cfg = dataclasses.replace(synth.PRESETS["colo829_shaped"], n_sv=6, coverage=7.0, read_len_mean=4000,
                              read_len_min=1500, read_len_max=9000, contig_len=150000, n_contigs=2, seed=78,
                              validate_sequence_length=600)
    wl = synth.make_workload(cfg)
    _, info, _ = _run_both(tmp_path, wl, "tumor", engine, L=600, min_mapq=0)
    assert len(info) >= 3


def test_insertion_heavy(engine, tmp_path):
    cfg = dataclasses.replace(synth.PRESETS["insertion_heavy"], n_sv=4, coverage=7.0, read_len_mean=5000,
                              read_len_min=2500, read_len_max=9000, contig_len=150000, n_contigs=2, seed=79,
                              validate_sequence_length=500, ins_len_range=(300, 1500))
    wl = synth.make_workload(cfg)
    _, info, _ = _run_both(tmp_path, wl, "tumor", engine, L=500, min_mapq=0, ratio=1.4)
    assert len(info) >= 2


def test_single_breakend(engine, tmp_path):
    cfg = dataclasses.replace(synth.PRESETS["single_breakend"], n_sv=10, coverage=8.0, read_len_mean=3000,
                              read_len_min=900, read_len_max=7000, contig_len=150000, n_contigs=2, seed=80,
                              sbnd_contig_range=(100, 900), sbnd_tail_range=(2000, 7000))
    wl = synth.make_workload(cfg)
    for sample in ("tumor", "control"):
        _, info, _ = _run_both(tmp_path, wl, sample, engine, is_sbnd=True)
        if sample == "tumor":
            assert len(info) >= 5


def test_all_alignment_tuples_of_all_reads(engine):
    """Not only the supporting reads: score, strand and end coordinates of every (read, template) and the begin
    coordinates wherever the device ran the begin pass must equal the oracle's ssw_check."""
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=6, coverage=8.0, read_len_mean=2000,
                              read_len_min=600, read_len_max=5000, contig_len=120000, n_contigs=2, seed=81,
                              short_fraction=0.7)
    wl = synth.make_workload(cfg)
    pack, svs, reads = wl.to_batch("tumor", var_read_min_mapq=0)
    plan = engine.plan(pack, svs, reads, prune=False)      # every alignment the reference runs
    plan.run()
    checked, supporting, records = plan.download()
    alns, flags = plan.download_alns()
    st = plan.stats()
    assert st["cells_executed"] == st["cells_dedup"] <= st["cells_reference"] and st["fwd_tasks"] > 0
    assert st["cells_pruned"] == 0 and st["reads_pruned"] == 0
    codes, offsets, lens = pack.finalize()

    def seq(i):
        return synth.codes_to_str(codes[int(offsets[i]):int(offsets[i]) + int(lens[i])])

    n_full = 0
    for r in range(len(reads)):
        sv = svs[reads["sv"][r]]
        rd = seq(reads["seq"][r])
        for s in range(4):
            if sv["tmpl"][s] == _lib.NMSV_NONE:
                assert alns[r, s]["strand"] == 0
                continue
            exp = O.ssw_check(seq(sv["tmpl"][s]), [("r", rd)], engine="striped")["r"]
            a = alns[r, s]
            strand = "+" if a["strand"] > 0 else "-"
            assert (int(a["score"]), strand, int(a["tend"])) == (exp[0], exp[5], exp[4])
            if flags[r] & 1:     # begin pass ran: the full 6-tuple is defined
                assert [int(a["score"]), int(a["qstart"]), int(a["qend"]), int(a["tstart"]), int(a["tend"]), strand] == exp
                n_full += 1
            elif strand == "+":
                assert int(a["qend"]) == exp[2]
            else:
                assert int(a["qstart"]) == exp[1]
    assert n_full > 10
    assert checked.tolist() == [int(e - b) for b, e in zip(svs["read_begin"], svs["read_end"])]
    assert int(supporting.sum()) == len(records) == int(((flags & 2) != 0).sum())
    # a plan can be re-run: identical results
    plan.run()
    c2, s2, r2 = plan.download()
    assert (s2 == supporting).all() and (r2 == records).all()
    plan.close()
    # the default plan skips alignments that cannot change an output: same counts and records, fewer cells; what it
    # did compute equals the full plan's tuples, a pruned read (flag 4) has none
    plan = engine.plan(pack, svs, reads)
    plan.run()
    c3, s3, r3 = plan.download()
    alns3, flags3 = plan.download_alns()
    st3 = plan.stats()
    plan.close()
    assert (c3 == checked).all() and (s3 == supporting).all() and (r3 == records).all()
    assert st3["cells_executed"] + st3["cells_pruned"] == st3["cells_dedup"] == st["cells_dedup"]
    assert 0 < st3["cells_executed"] < st["cells_executed"] and st3["reads_pruned"] > 0
    assert st3["reads_pruned"] == int(((flags3 & 4) != 0).sum())
    assert ((flags3 & 3) == (flags & 3)).all()
    for r in range(len(reads)):
        for s in range(4):
            if alns3[r, s]["strand"] != 0:
                assert alns3[r, s] == alns[r, s]
            if flags3[r] & 4:
                assert alns3[r, s]["strand"] == 0 and alns3[r, s]["score"] == 0


def test_edge_cases(engine, tmp_path):
    """Empty input, SV without reads (no output line), duplicate read ids (last wins), empty read sequence (dropped),
    mapq None, reads shorter than the template."""
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=3, coverage=5.0, read_len_mean=1500,
                              read_len_min=500, read_len_max=3000, contig_len=80000, n_contigs=2, seed=82)
    wl = synth.make_workload(cfg)
    seam = str(tmp_path / "edge.tsv")
    wl.write_seam("tumor", seam)
    lines = open(seam).read().splitlines()
    k0 = lines[0].split("\t")[0]
    first = lines[0].split("\t")
    extra = [
        "\t".join([k0, first[1], "60", "60", first[4][:150]]),          # duplicate id, later (shorter) sequence wins
        "\t".join([k0, "emptyread", "60", "60", ""]),                   # dropped by the FASTA reader semantics
        "\t".join([k0, "tiny", "None", "60", "ACGTACGTAC"]),             # shorter than any template, mapq None
        "\t".join([k0, "allN", "60", "60", "N" * 700]),
    ]
    rows = [ln for ln in lines if ln.split("\t")[0] == k0] + extra + [ln for ln in lines if ln.split("\t")[0] != k0]
    open(seam, "w").write("\n".join(rows) + "\n")
    O.count_from_seam(seam, wl.fasta, seam + ".o_count", seam + ".o_info", var_read_min_mapq=0)
    C.count_sread_from_seam(seam, seam + ".g_count", seam + ".g_info", wl.fasta, False, 0, 1.2, engine=engine)
    assert open(seam + ".g_count").read() == open(seam + ".o_count").read()
    assert sorted(open(seam + ".g_info").read().splitlines()) == sorted(open(seam + ".o_info").read().splitlines())
    # empty seam file -> empty outputs
    empty = str(tmp_path / "empty.tsv")
    open(empty, "w").close()
    C.count_sread_from_seam(empty, empty + ".count", empty + ".info", wl.fasta, False, 0, 1.2, engine=engine)
    assert open(empty + ".count").read() == "" and open(empty + ".info").read() == ""
    # empty batch through the C ABI
    checked, supporting, records = engine.validate(*synth.make_workload(dataclasses.replace(cfg, n_sv=0)).to_batch("tumor"))
    assert len(checked) == 0 and len(records) == 0


def test_batches_are_split_and_merged_transparently(engine, tmp_path, monkeypatch):
    cfg = dataclasses.replace(synth.PRESETS["testdata_shaped"], n_sv=9, coverage=5.0, read_len_mean=1500,
                              read_len_min=500, read_len_max=3000, contig_len=80000, n_contigs=2, seed=83)
    wl = synth.make_workload(cfg)
    monkeypatch.setattr(C, "BATCH_BASES", 20000)      # forces several device batches
    _, _, counter = _run_both(tmp_path, wl, "tumor", engine)
    assert counter.batches_run >= 3


def test_ssw_check_file_seam(engine, tmp_path):
    q = tmp_path / "var.fa"
    t = tmp_path / "reads.fa"
    q.write_text(">k\nACGTTTTTACGT\n")
    t.write_text(">r1 extra\nACGTACGT\n>r2\nAAAATTTT\n>r1\nACGTTTTTACGTAA\n")
    got = C.ssw_check(str(q), str(t), False, engine=engine)
    exp = O.ssw_check("ACGTTTTTACGT", list(O.fasta_records(str(t))))
    assert got == exp and set(got) == {"r1", "r2"}


def test_fused_and_separate_begin_pass_agree(engine, monkeypatch):
    """By default the forward launches also drain the begin-pass queue that the per-read continuation fills
    (DESIGN.md "fused begin pass"); NMSV_FUSED_BEGIN=0 runs decide and the begin pass as launches of their own. Both
    must produce identical counts, records and alignment tuples -- on a canonical workload (one strip class) and on a
    single-breakend one (two strip classes: 601- and 400-base templates), and on repeated runs of one plan."""
    for preset, kw in (("testdata_shaped", dict(n_sv=10, coverage=10.0, short_fraction=0.5, seed=311)),
                       ("single_breakend", dict(n_sv=8, coverage=10.0, seed=312))):
        cfg = dataclasses.replace(synth.PRESETS[preset], read_len_mean=2500, read_len_min=500, read_len_max=8000,
                                  contig_len=150000, n_contigs=2, **kw)
        wl = synth.make_workload(cfg)
        pack, svs, reads = wl.to_batch("tumor", var_read_min_mapq=0)
        out = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("NMSV_FUSED_BEGIN", mode)
            for prune in (False, True):
                plan = engine.plan(pack, svs, reads, prune=prune)
                for _ in range(2):
                    plan.run()
                checked, supporting, records = plan.download()
                alns, flags = plan.download_alns()
                st = plan.stats()
                assert st["kernel_launches"] == (2 + 5 * st["n_classes"] if mode == "1" else 3 + 2 * st["n_classes"])
                out[mode, prune] = (checked, supporting, records, alns, flags, st["rev_tasks"])
                plan.close()
        a, b = out["1", False], out["0", False]
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (a[2] == b[2]).all()
        assert (a[3] == b[3]).all() and (a[4] == b[4]).all() and a[5] == b[5] > 0
        assert int(a[1].sum()) > 0
        for key in (("1", True), ("0", True)):      # pruned plans: same counts, records and begin tasks
            c = out[key]
            assert (a[0] == c[0]).all() and (a[1] == c[1]).all() and (a[2] == c[2]).all() and a[5] == c[5]


def test_pruning_never_changes_an_output(engine):
    """nmsv_set_option("prune"): reads whose mapq test fails are only counted and ref1/ref2 are aligned only for reads
    that pass the variant tests (reference :336-346, :384-389). Counts and records must equal the unpruned plan's on
    every workload shape, with and without a mapq threshold."""
    shapes = (("testdata_shaped", dict(n_sv=12, coverage=9.0, short_fraction=0.7, seed=411), 200),
              ("colo829_shaped", dict(n_sv=6, coverage=7.0, seed=412, validate_sequence_length=600), 600),
              ("single_breakend", dict(n_sv=10, coverage=9.0, seed=413, sbnd_contig_range=(100, 900),
                                       sbnd_tail_range=(2000, 7000)), 200))
    for preset, kw, L in shapes:
        cfg = dataclasses.replace(synth.PRESETS[preset], read_len_mean=2500, read_len_min=600, read_len_max=7000,
                                  contig_len=150000, n_contigs=2, **kw)
        wl = synth.make_workload(cfg)
        for sample in ("tumor", "control"):
            for min_mapq in (0, 40):
                pack, svs, reads = wl.to_batch(sample, var_read_min_mapq=min_mapq)
                res = {}
                for prune in (False, True):
                    plan = engine.plan(pack, svs, reads, prune=prune)
                    plan.run()
                    res[prune] = plan.download() + (plan.stats(),)
                    plan.close()
                full, fast = res[False], res[True]
                assert (full[0] == fast[0]).all() and (full[1] == fast[1]).all() and (full[2] == fast[2]).all()
                assert fast[3]["cells_executed"] <= full[3]["cells_executed"]
                if sample == "tumor":
                    assert int(full[1].sum()) > 0


def test_abandoned_reference_alleles_leave_the_rings_consistent(engine):
    """Short deletions / duplications make most reads pass the variant tests and fail the margin test: their reference-allele
    tasks are abandoned mid-way (as soon as a score comes within the margin), at every phase of a chunk and of the padded
    last steps. Whatever an abandoned task leaves in the boundary rings must not reach later tasks: many such tasks in one
    plan, run several times on the same context, must give the counts and records of a plan that computes everything."""
    cfg = dataclasses.replace(synth.PRESETS["colo829_shaped"], n_sv=40, coverage=12.0, read_len_mean=2500, read_len_min=600,
                              read_len_max=9000, contig_len=300000, n_contigs=2, seed=517, validate_sequence_length=700,
                              short_fraction=0.95, sv_mix=(("DEL", 0.6), ("DUP", 0.4)))
    wl = synth.make_workload(cfg)
    for sample in ("tumor", "control"):
        pack, svs, reads = wl.to_batch(sample, var_read_min_mapq=0)
        plan = engine.plan(pack, svs, reads, prune=False)
        plan.run()
        full = plan.download()
        plan.close()
        plan = engine.plan(pack, svs, reads)
        for _ in range(4):
            plan.run()
            got = plan.download()
            assert (got[0] == full[0]).all() and (got[1] == full[1]).all() and (got[2] == full[2]).all()
        st = plan.stats()
        plan.close()
        assert st["fwd_tasks_queued"] > 200 and st["cells_executed"] < 0.8 * st["cells_dedup"]
    assert int(full[1].sum()) >= 0
