"""CPU tests: the C-ABI library builds, loads and exports what include/nmsv.h declares; host logic."""
import ctypes
import math
import os
import re

import numpy as np
import pytest

import nanomonsv_b200
from nanomonsv_b200 import _lib, batch, count_sread_by_alignment as C, my_seq, synth
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "nmsv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nmsv_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    declared = _header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/nmsv.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert L.nmsv_abi_version() == _lib.ABI_VERSION == 2


def test_struct_layouts_match_header():
    assert _lib.SV_DTYPE.itemsize == 80
    assert _lib.READ_DTYPE.itemsize == 16
    assert _lib.ALN_DTYPE.itemsize == 24
    assert _lib.RECORD_DTYPE.itemsize == 104
    assert _lib.SSW_RESULT_DTYPE.itemsize == 20
    assert ctypes.sizeof(_lib.Stats) == 112


def test_encode_ascii_matches_python_lut():
    L = _lib.load()
    s = bytes(range(1, 256))
    out = np.zeros(len(s), dtype=np.uint8)
    L.nmsv_encode_ascii(s, len(s), out.ctypes.data)
    assert (out == my_seq.CODE_LUT[np.frombuffer(s, dtype=np.uint8)]).all()
    assert (out == O.encode(s.decode("latin-1"))).all()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nanomonsv_b200.EngineError) as e:
        nanomonsv_b200.Engine(0)
    assert "no CPU fallback" in str(e.value)


def test_integer_bounds_equal_float_comparisons():
    for ratio in (1.2, 1.4, 1.6, 1.8):
        for length in list(range(1, 700)) + [2000, 2001, 6000]:
            smin, qmax, emin = batch.integer_bounds(length, ratio, 0.1, 0.9)
            for score in (smin - 1, smin):
                assert (score > ratio * length) == (score >= smin)
            for q in (qmax, qmax + 1):
                assert (q < 0.1 * length) == (q <= qmax)
            for q in (emin - 1, emin):
                assert (q > 0.9 * length) == (q >= emin)
    assert batch.integer_bounds(401, 1.8, 0.1, 0.9)[0] == math.floor(721.8000000000001) + 1


def test_reverse_complement_matches_oracle():
    rng = np.random.default_rng(3)
    alpha = "ACGTNacgtnWSMKRYBDHVxX-"
    for _ in range(500):
        s = "".join(alpha[i] for i in rng.integers(0, len(alpha), size=int(rng.integers(0, 50))))
        assert my_seq.reverse_complement(s) == O.reverse_complement(s)
    assert my_seq.needs_explicit_revcomp("ACGa") and not my_seq.needs_explicit_revcomp("ACGTN")


def test_templates_match_oracle_restatement():
    cfg = synth.SynthConfig(n_sv=40, contig_len=120000, n_contigs=3, read_len_max=3000, read_len_mean=1000,
                            read_len_min=500, coverage=0.0, seed=5, short_fraction=0.5)
    wl = synth.make_workload(cfg, samples=())
    fasta = wl.fasta
    n_short = 0
    for sv in wl.svs:
        for L in (200, 1000):
            assert C.canonical_templates(sv.key, fasta, L) == O.sv_templates(sv.key, fasta, L)
        n_short += C.canonical_templates(sv.key, fasta, 200)[4]
    assert 0 < n_short < len(wl.svs)
    # near a contig start the left flank is clipped (reference :239: max(pos - L - 1, 0))
    key = "chr1,50,+,chr2,1000,-,---,r_0"
    v1, v2, r1, r2, short = C.canonical_templates(key, fasta, 200)
    assert (v1, v2, r1, r2, short) == O.sv_templates(key, fasta, 200)
    assert len(r1) == 249 and len(v1) == 49 + 200
    # the is_short_del_dup quirk: span + len(str(len(id))) < 300, insertion length ignored (:210-223)
    assert C.canonical_templates("chr1,1000,+,chr1,1298,-,---,r_1", fasta, 200)[4] is True     # 298 + len("0") = 299
    assert C.canonical_templates("chr1,1000,+,chr1,1299,-,---,r_1", fasta, 200)[4] is False
    assert C.canonical_templates("chr1,1000,+,chr1,1298,-,ACGT,r_1", fasta, 200)[4] is True    # len(str(len("r_1"))) = 1
    assert C.canonical_templates("chr1,1000,+,chr1,1298,-,ACGT,r_123456789", fasta, 200)[4] is False  # len("11") = 2
    sb = synth.make_workload(synth.SynthConfig(n_sv=10, sbnd=True, contig_len=150000, n_contigs=2, read_len_max=3000,
                                               coverage=0.0, seed=6, validate_sequence_length=200), samples=())
    for sv in sb.svs:
        var1, ref1 = C.sbnd_templates(sv.key, sb.fasta, 200)
        assert (var1, ref1) == O.sbnd_templates(sv.key, sb.fasta, 200)
        assert len(var1) in (601, 400) or len(sv.key.split(",")[3]) < 200


def test_read_fasta_semantics(tmp_path):
    p = tmp_path / "r.fa"
    p.write_text(">a desc\nACGT\nAC\n>empty\n>b\nTT\n>last\n")
    assert list(C.read_fasta(str(p))) == list(O.fasta_records(str(p))) == [("a", "ACGTAC"), ("b", "TT"), ("last", "")]


def test_batch_builder_layout():
    bb = batch.BatchBuilder(score_ratio_thres=1.2, var_read_min_mapq=40)
    bb.add_sv(["ACGT" * 100, "ACGT" * 100, None, None], False, False)
    bb.add_read("ACGT" * 300, 60, None)
    bb.add_read(np.zeros(50, dtype=np.uint8), 3, 60)
    bb.add_sv(["ACGa" * 50, None, "TTTT" * 50, None], True, False)
    pack, svs, reads = bb.finalize()
    codes, offsets, lens = pack.finalize()
    assert svs["tmpl"][0][0] == svs["tmpl"][0][1] and svs["tmpl"][0][2] == _lib.NMSV_NONE
    assert tuple(svs["score_min"][0]) == (481, 481) and tuple(svs["qstart_max"][0]) == (39, 39)
    assert (svs["read_begin"][0], svs["read_end"][0]) == (0, 2) and (svs["read_begin"][1], svs["read_end"][1]) == (2, 2)
    assert svs["tmpl_rc"][1][0] != _lib.NMSV_NONE and svs["flags"][1] == _lib.SV_SBND
    assert reads["mapq2"][0] == _lib.MAPQ_NONE and reads["sv"].tolist() == [0, 0]
    # sequences start on 16-byte boundaries (TMA staging granularity), gaps hold the wildcard code
    assert (offsets % 16 == 0).all() and codes.size == int(((lens.astype(np.int64) + 15) // 16 * 16).sum())
    assert (codes[int(offsets[0]) + int(lens[0]):int(offsets[1])] == 4).all()


def test_synth_workload_is_deterministic_and_shaped():
    cfg = synth.SynthConfig(n_sv=12, contig_len=150000, n_contigs=2, read_len_mean=2000, read_len_min=800,
                            read_len_max=5000, coverage=10.0, seed=9, validate_sequence_length=200)
    a, b = synth.make_workload(cfg), synth.make_workload(cfg)
    assert [s.key for s in a.svs] == [s.key for s in b.svs]
    for sample in ("tumor", "control"):
        for ra, rb in zip(a.samples[sample], b.samples[sample]):
            assert [r.rid for r in ra] == [r.rid for r in rb]
            assert all((x.codes == y.codes).all() for x, y in zip(ra, rb))
    n_t = sum(len(x) for x in a.samples["tumor"])
    assert 12 * 10 <= n_t <= 12 * 30
    pack, svs, reads = a.to_batch("tumor")
    assert len(svs) == 12 and len(reads) == n_t


class _RecordingEngine:
    """Stands in for the GPU engine: keeps every batch it is handed (host logic only)."""

    def __init__(self):
        self.batches = []

    def validate(self, pack, svs, reads, scoring=None):
        codes, offsets, lens = pack.finalize()
        self.batches.append((codes.copy(), offsets.copy(), lens.copy(), svs.copy(), reads.copy()))
        checked = np.array([int(e) - int(b) for b, e in zip(svs["read_begin"], svs["read_end"])], dtype=np.int32)
        return checked, np.zeros(len(svs), dtype=np.int32), np.zeros(0, dtype=_lib.RECORD_DTYPE)


@pytest.mark.parametrize("sbnd", [False, True])
def test_byte_level_seam_reader_builds_the_same_batches(tmp_path, monkeypatch, sbnd):
    """The block / memchr / nmsv_encode_pack reader (default) against the line-by-line reader that mirrors the
    reference's loop (:414-425): identical code buffers, SV and read tables, count files -- including CR LF line ends,
    a last line without newline, repeated read ids (last record wins), empty sequences (mapq kept, read dropped),
    'None' mapqs, soft-masked and N bases, one batch and several batches."""
    cfg = synth.PRESETS["single_breakend" if sbnd else "testdata_shaped"]
    import dataclasses
    cfg = dataclasses.replace(cfg, n_sv=6, coverage=6.0, read_len_mean=900, read_len_min=300, read_len_max=2000,
                              contig_len=80000, n_contigs=2, seed=5, **({} if sbnd else {"short_fraction": 0.5}))
    wl = synth.make_workload(cfg, samples=("tumor",))
    seam = str(tmp_path / "seam.tsv")
    wl.write_seam("tumor", seam)
    lines = open(seam).read().splitlines()
    k0 = lines[0].split("\t")[0]
    extra = [f"{k0}\tdup_read\t60\t60\tACGTNNacgt" + "ACGT" * 50, f"{k0}\tdup_read\tNone\t30\t" + "TTGCA" * 40,   # same id twice
             f"{k0}\tempty_read\t60\t60\t", f"{k0}\tspaced id tail\t12\tNone\t  " + "GATTACA" * 30 + " "]
    body = "\r\n".join(extra + lines)                        # CR LF everywhere, no newline after the last line
    open(seam, "w", newline="").write(body)
    out = {}
    for reader, batch_bases in (("lines", 512 << 20), ("bytes", 512 << 20), ("lines", 20000), ("bytes", 20000)):
        monkeypatch.setenv("NMSV_SEAM_READER", reader)
        monkeypatch.setattr(C, "BATCH_BASES", batch_bases)       # 20000: several batches in flight behind the parser
        eng = _RecordingEngine()
        fc, fi = str(tmp_path / f"{reader}{batch_bases}.count"), str(tmp_path / f"{reader}{batch_bases}.info")
        C.count_sread_from_seam(seam, fc, fi, wl.fasta, sbnd, 40, 1.2, engine=eng)
        out[(reader, batch_bases)] = (eng.batches, open(fc).read())
    ref_batches, ref_count = out[("lines", 512 << 20)]
    assert len(ref_batches) == 1 and ref_count.count("\n") == 6 and len(out[("lines", 20000)][0]) > 1
    for key, ref_key in ((("bytes", 512 << 20), ("lines", 512 << 20)), (("bytes", 20000), ("lines", 20000))):
        batches, count = out[key]
        ref_batches, ref_count = out[ref_key]
        assert count == ref_count and len(batches) == len(ref_batches)
        for a, b in zip(ref_batches, batches):
            for x, y in zip(a, b):
                assert x.dtype == y.dtype and x.shape == y.shape and (x == y).all(), key
    # the library's one-pass scan (nmsv_seam_scan, the default reader) lays a batch out differently (templates first, one
    # table per batch, mapq-failing reads as lengths only): same candidates, templates, reads in the same order, counts
    def canonical(batches):
        res = []
        for codes, offsets, lens, svs, reads in batches:
            def seq(i):
                if int(offsets[i]) == _lib.SEQ_ABSENT:
                    return ("absent", int(lens[i]))
                return bytes(codes[int(offsets[i]):int(offsets[i]) + int(lens[i])])
            for sv in svs:
                tm = tuple(seq(int(t)) if int(t) != _lib.NMSV_NONE else None for t in sv["tmpl"])
                rc = tuple(seq(int(t)) if int(t) != _lib.NMSV_NONE else None for t in sv["tmpl_rc"])
                rs = [(int(r["mapq1"]), int(r["mapq2"]), seq(int(r["seq"]))) for r in reads[int(sv["read_begin"]):int(sv["read_end"])]]
                res.append((tm, rc, rs, tuple(int(x) for x in sv["score_min"]), tuple(int(x) for x in sv["qstart_max"]),
                            tuple(int(x) for x in sv["qend_min"]), int(sv["margin"]), int(sv["min_mapq"]), int(sv["flags"])))
        return res
    for batch_bases in (512 << 20, 20000):
        monkeypatch.setenv("NMSV_SEAM_READER", "native")
        monkeypatch.setattr(C, "BATCH_BASES", batch_bases)
        eng = _RecordingEngine()
        fc, fi = str(tmp_path / f"native{batch_bases}.count"), str(tmp_path / f"native{batch_bases}.info")
        C.count_sread_from_seam(seam, fc, fi, wl.fasta, sbnd, 40, 1.2, engine=eng)
        ref_batches, ref_count = out[("lines", batch_bases)]
        assert open(fc).read() == ref_count       # (the scan starts with smaller batches: their number may differ)
        assert (len(eng.batches) > 1) == (len(ref_batches) > 1)
        assert canonical(eng.batches) == canonical(ref_batches)
        assert any(int(o) == _lib.SEQ_ABSENT for b in eng.batches for o in b[1])
    # the extra lines did what they are there for
    codes, offsets, lens, svs, reads = out[("lines", 512 << 20)][0][0]
    n_first = int(svs["read_end"][0]) - int(svs["read_begin"][0])
    assert n_first == sum(1 for ln in lines if ln.split("\t")[0] == k0) + 2   # dup_read once, empty_read dropped
    open(seam, "a").write("\nbroken line without tabs\n")
    for reader in ("bytes", "native"):
        monkeypatch.setenv("NMSV_SEAM_READER", reader)
        with pytest.raises(ValueError):
            C.count_sread_from_seam(seam, fc, fi, wl.fasta, sbnd, 40, 1.2, engine=_RecordingEngine())


def test_device_pick_follows_env_then_pool_worker(monkeypatch):
    """default_engine(): NMSV_DEVICE, then LOCAL_RANK, then the multiprocessing-pool worker number modulo the number of
    GPUs (the reference's --processes pool, run.py:361-372), then 0."""
    import multiprocessing
    from nanomonsv_b200 import engine
    monkeypatch.delenv("NMSV_DEVICE", raising=False)
    monkeypatch.delenv("LOCAL_RANK", raising=False)
    assert engine.pick_device() == 0
    monkeypatch.setenv("LOCAL_RANK", "3")
    assert engine.pick_device() == 3
    monkeypatch.setenv("NMSV_DEVICE", "5")
    assert engine.pick_device() == 5
    monkeypatch.delenv("NMSV_DEVICE")
    monkeypatch.delenv("LOCAL_RANK")

    class FakeLib:
        def nmsv_device_count(self):
            return 8
    monkeypatch.setattr(engine._lib, "load", lambda: FakeLib())
    monkeypatch.setattr(multiprocessing.current_process(), "_identity", (11,), raising=False)
    assert engine.pick_device() == (11 - 1) % 8


def test_window_subsampling_equals_the_reference_draws():
    """bam_subsample_fetch (reference :11-23): above max_read_num reads in a window a random subset is kept. The reference
    draws unseeded; with the same seed the mirror must keep exactly the reads the reference's own function kept
    (tests/golden/subsample_vectors.json, incl. the > 500-read branch), in file order."""
    import json
    import random
    vec = json.load(open(os.path.join(ROOT, "tests", "golden", "subsample_vectors.json")))["cases"]

    class Handle:
        def __init__(self, n):
            self.n = n

        def fetch(self, chrom, start, end):
            return iter(range(self.n))
    assert any(c["n"] > 500 and len(c["kept"]) == 500 for c in vec)
    for c in vec:
        random.seed(c["seed"])
        assert list(C.bam_subsample_fetch(Handle(c["n"]), "chr1", 0, 200, c["max_read_num"])) == c["kept"], c["n"]


def _scan_all(data: bytes, max_bases, tail_bytes, is_sbnd, gcap, rcap):
    """Drive nmsv_seam_scan over a whole buffer the way the stage function does; returns [(key, [(id, seq, m1, m2)])]."""
    L = _lib.load()
    src = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
    groups = np.zeros(gcap, dtype=_lib.SEAM_GROUP_DTYPE)
    reads = np.zeros(rcap, dtype=_lib.SEAM_READ_DTYPE)
    ng, nr, nxt, err = ctypes.c_uint32(0), ctypes.c_uint32(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
    pos, out, calls = 0, [], 0
    while pos < len(data):
        rc = L.nmsv_seam_scan(src.ctypes.data, len(data), pos, max_bases, tail_bytes, is_sbnd, groups.ctypes.data, gcap,
                              reads.ctypes.data, rcap, ctypes.byref(ng), ctypes.byref(nr), ctypes.byref(nxt), ctypes.byref(err))
        if rc != 0:
            return rc, err.value
        assert nxt.value > pos and ng.value >= 1
        for g in groups[:ng.value]:
            rows = []
            for r in reads[int(g["read_begin"]):int(g["read_end"])]:
                rows.append((data[int(r["id_off"]):int(r["id_off"]) + int(r["id_len"])].decode(),
                             data[int(r["seq_off"]):int(r["seq_off"]) + int(r["seq_len"])].decode(), int(r["mapq1"]), int(r["mapq2"])))
            out.append((data[int(g["key_off"]):int(g["key_off"]) + int(g["key_len"])].decode(), rows))
        pos = nxt.value
        calls += 1
    return out, calls


def test_seam_scan_grouping_capacity_and_errors():
    """nmsv_seam_scan: whole key groups per call, the dict semantics of the reference for repeated read ids, batch cuts by
    bases, by table capacity and by the short-rest rule, error offsets -- against a Python restatement of the loop."""
    import random
    rnd = random.Random(11)
    lines, expect = [], []
    for k in range(23):
        key = f"chr{k % 3},{1000 + k},+,chr1,{5000 + k},-,---,r_{k}"
        order, seqs, mq = [], {}, {}
        for r in range(rnd.randint(1, 9)):
            rid = f"read{rnd.randint(0, 5)}"
            seq = "" if rnd.random() < 0.15 else "".join(rnd.choice("ACGTNacgt") for _ in range(rnd.randint(1, 60)))
            m1 = "None" if rnd.random() < 0.2 else str(rnd.randint(0, 60))
            m2 = "None" if rnd.random() < 0.2 else str(rnd.randint(0, 60))
            lines.append(f"{key}\t{rid} extra words\t{m1}\t{m2}\t {seq} ")
            if rid in seqs or seq:
                mq[rid] = (m1, m2)
            if seq:
                if rid not in seqs:
                    order.append(rid)
                seqs[rid] = seq
        conv = lambda v: -1 if v == "None" else int(v)   # noqa: E731
        if order:
            expect.append((key, [(rid, seqs[rid], conv(mq[rid][0]), conv(mq[rid][1])) for rid in order]))
    data = ("\r\n".join(lines)).encode()                 # CR LF, no newline at the end
    ref, _ = _scan_all(data, 1 << 40, 0, 0, 4096, 4096)
    assert [g for g in ref if g[1]] == expect and len(ref) == 23      # a group whose reads are all empty still exists (0 reads)
    # cut by bases: several calls, same content; by tiny tables: same content; short rest joins the last batch
    for args in ((200, 0, 64, 64), (1 << 40, 0, 2, 4096), (1 << 40, 0, 4096, 12), (150, 10 ** 9, 4096, 4096)):
        got, calls = _scan_all(data, args[0], args[1], 0, args[2], args[3])
        assert got == ref, args
        assert calls == 1 if args[1] else calls > 1
    # single-breakend files: the fourth column is ignored (None)
    sb, _ = _scan_all(data, 1 << 40, 0, 1, 4096, 4096)
    assert all(r[3] == -1 for g in sb for r in g[1]) and [(g[0], [r[:3] for r in g[1]]) for g in sb] == [(g[0], [r[:3] for r in g[1]]) for g in ref]
    # one group larger than the read table: NMSV_E_CAPACITY (the caller grows the table)
    assert _scan_all(data, 1 << 40, 0, 0, 4096, 1)[0] == _lib.E_CAPACITY
    # malformed lines: NMSV_E_INVALID with the offset of the line
    bad = data + b"\nkey\tid\tnot_a_number\t3\tACGT\n"
    rc, off = _scan_all(bad, 1 << 40, 0, 0, 4096, 4096)
    assert rc == _lib.E_INVALID and bad[off:off + 3] == b"key"
    rc, off = _scan_all(b"only\tfour\t1\t2\n", 1 << 40, 0, 0, 16, 16)
    assert rc == _lib.E_INVALID and off == 0
    assert _scan_all(b"", 1, 0, 0, 4, 4) == ([], 0)


def test_sw_jump_strings_reproduce_the_reference_rows():
    """nmsv_sw_jump_strings (host side of the sw_jump row, no GPU work): fed with the traceback operations of the C oracle it
    must write the two alignment rows of the reference's own return values (tests/golden/sw_jump_vectors.json, produced by
    nanomonsv/smith_waterman.py:5-127 itself)."""
    import json
    L = _lib.load()
    OL = O.lib()
    vec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sw_jump_vectors.json")))["vectors"]
    seqs, results, ops_segments = [], [], []
    for v in vec:
        c, a, b = (np.frombuffer(v[k].encode("latin-1"), dtype=np.uint8) for k in ("contig", "region1", "region2"))
        with np.errstate(divide="ignore"):
            logtab = np.log(np.arange(len(c) + 1))
        res = O.JumpResult()
        ops = np.zeros(2 * len(c) + len(a) + len(b) + 4, dtype=np.uint8)
        assert OL.ora_sw_jump(c.ctypes.data, len(c), a.ctypes.data, len(a), b.ctypes.data, len(b), *v["params"],
                              logtab.ctypes.data, ctypes.byref(res), ops.ctypes.data) == 0
        seqs += [c, a, b]
        results.append(tuple(getattr(res, n) for n, _ in O.JumpResult._fields_))
        ops_segments.append(ops)
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    offsets = np.zeros(len(seqs), dtype=np.uint64)
    np.cumsum(lens[:-1].astype(np.uint64), out=offsets[1:])
    data = np.concatenate(seqs)
    n = len(vec)
    idx = np.arange(3 * n, dtype=np.uint32)
    cidx, aidx, bidx = (np.ascontiguousarray(idx[k::3]) for k in range(3))
    out = np.array(results, dtype=_lib.JUMP_RESULT_DTYPE)
    ops = np.concatenate(ops_segments)
    row_off = np.zeros(n, dtype=np.uint64)
    np.cumsum(np.array([len(o) for o in ops_segments[:-1]], dtype=np.uint64), out=row_off[1:])
    crow, rrow = np.zeros(ops.size, dtype=np.uint8), np.zeros(ops.size, dtype=np.uint8)
    assert L.nmsv_sw_jump_strings(data.ctypes.data, offsets.ctypes.data, lens.ctypes.data, cidx.ctypes.data, aidx.ctypes.data,
                                  bidx.ctypes.data, n, out.ctypes.data, ops.ctypes.data, row_off.ctypes.data,
                                  crow.ctypes.data, rrow.ctypes.data) == 0
    checked = 0
    for k, v in enumerate(vec):
        if v["expected"] is None:
            assert not out[k]["found"]
            continue
        a = int(row_off[k]); b = a + int(out[k]["n_ops1"]) + 1 + int(out[k]["n_ops2"])
        assert crow[a:b].tobytes().decode("latin-1") == v["expected"][4]
        assert rrow[a:b].tobytes().decode("latin-1") == v["expected"][5]
        checked += 1
    assert checked > 90
    # operations that do not belong to the sequences are refused, not read out of bounds
    bad = ops.copy()
    k = next(i for i, v in enumerate(vec) if v["expected"] is not None)
    bad[int(row_off[k]):int(row_off[k]) + int(out[k]["n_ops2"])] = 2       # walks the contig far beyond its start
    big = out.copy(); big[k]["n_ops2"] = 2 * lens[3 * k] + lens[3 * k + 2] + 2; big[k]["n_ops1"] = 1
    bad[int(row_off[k]):int(row_off[k]) + int(big[k]["n_ops2"]) + 1] = 2
    assert L.nmsv_sw_jump_strings(data.ctypes.data, offsets.ctypes.data, lens.ctypes.data, cidx.ctypes.data, aidx.ctypes.data,
                                  bidx.ctypes.data, n, big.ctypes.data, bad.ctypes.data, row_off.ctypes.data,
                                  crow.ctypes.data, rrow.ctypes.data) == _lib.E_INVALID
