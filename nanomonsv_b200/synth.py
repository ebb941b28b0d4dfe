"""Synthetic ONT-like workloads shaped like the BASELINE.json configs.

Everything is generated in code space (uint8: 0..3 = ACGT, 4 = N) with numpy and a fixed seed; the same workload can
be written as the reference's seam TSV (`key \\t qname \\t mapq1 \\t mapq2 \\t sequence`, count_sread_by_alignment.py
:111,:120) for the oracle and the file-level drop-in, or packed straight into a device batch for bench.py.

There is no network and no real data in this environment; the test BAMs named by BASELINE config 1 are absent
from the reference checkout (.MISSING_LARGE_BLOBS), so config 1 is represented by a stand-in of the same shape.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .batch import BatchBuilder
from .count_sread_by_alignment import canonical_templates, sbnd_templates

_ASCII = np.frombuffer(b"ACGTN", dtype=np.uint8)
_COMP = np.array([3, 2, 1, 0, 4], dtype=np.uint8)


def codes_to_str(codes: np.ndarray) -> str:
    return _ASCII[codes].tobytes().decode("ascii")


def revcomp_codes(codes: np.ndarray) -> np.ndarray:
    return _COMP[codes[::-1]]


class CodeFasta:
    """fetch(chrom, start0, end0) over in-memory contigs, like pysam.FastaFile.fetch (half open, clipped)."""

    def __init__(self, contigs: Dict[str, np.ndarray]):
        self.contigs = contigs

    def fetch(self, chrom: str, start: int, end: int) -> str:
        c = self.contigs[chrom]
        return codes_to_str(c[max(start, 0):max(end, 0)])

    def get_reference_length(self, chrom: str) -> int:
        return len(self.contigs[chrom])

    def close(self) -> None:
        pass


@dataclass
class SynthConfig:
    name: str = "colo829_shaped"
    n_sv: int = 100
    validate_sequence_length: int = 1000       # template length = 2 * this (reference default 200 -> 400)
    coverage: float = 30.0
    tumor_vaf: float = 0.3
    control_vaf: float = 0.0
    read_len_mean: float = 10000.0
    read_len_shape: float = 2.0
    read_len_min: int = 1000
    read_len_max: int = 50000
    sub_rate: float = 0.03
    ins_rate: float = 0.025
    del_rate: float = 0.025
    sv_mix: Tuple[Tuple[str, float], ...] = (("DEL", 0.40), ("INS", 0.25), ("DUP", 0.10), ("INV", 0.10), ("TRA", 0.15))
    short_fraction: float = 0.15               # share of DEL/DUP with a span < 300 (the 8-alignment branch)
    ins_len_range: Tuple[int, int] = (50, 500)
    n_contigs: int = 4
    contig_len: int = 2_000_000
    n_run_rate: float = 0.001                  # fraction of reference bases inside short N runs
    low_mapq_fraction: float = 0.05
    sbnd: bool = False
    sbnd_contig_range: Tuple[int, int] = (200, 5000)
    sbnd_tail_range: Tuple[int, int] = (5000, 50000)
    stratified: bool = False                   # SV kinds (and the short DEL/DUP share) in exact proportions per 40 candidates
                                               # instead of independent draws: batches of a few hundred candidates then carry
                                               # the same mix on every rank (bench.py: weak scaling compares like with like)
    seed: int = 20261019


PRESETS: Dict[str, SynthConfig] = {
    # config 1 stand-in: repo test data shape (m = 400, long reads 5-39 kb, ~46 tumor / ~37 control reads)
    "testdata_shaped": SynthConfig(name="testdata_shaped", n_sv=25, validate_sequence_length=200, coverage=40.0,
                                   read_len_mean=15000.0, read_len_min=5000, read_len_max=39000, seed=20261019 + 1),
    # config 2: COLO829-shaped, 2 kb templates
    "colo829_shaped": SynthConfig(name="colo829_shaped", n_sv=20000, validate_sequence_length=1000,
                                  seed=20261019 + 2),
    # config 3: insertion heavy, 6 kb templates
    "insertion_heavy": SynthConfig(name="insertion_heavy", n_sv=5000, validate_sequence_length=3000, coverage=40.0,
                                   sv_mix=(("INS", 1.0),), ins_len_range=(300, 6000), seed=20261019 + 3),
    # config 4: single breakends with long soft-clipped tails
    "single_breakend": SynthConfig(name="single_breakend", n_sv=5000, validate_sequence_length=200, sbnd=True,
                                   seed=20261019 + 4),
}


@dataclass
class SynthSV:
    key: str
    is_sbnd: bool
    kind: str


@dataclass
class SynthRead:
    rid: str
    mapq1: Optional[int]
    mapq2: Optional[int]
    codes: np.ndarray


@dataclass
class Workload:
    config: SynthConfig
    fasta: CodeFasta
    svs: List[SynthSV]
    samples: Dict[str, List[List[SynthRead]]] = field(default_factory=dict)   # sample -> per SV -> reads

    # ---- the reference's seam TSV, key-sorted (plain byte order, i.e. LC_ALL=C sort -k1,1)
    def write_seam(self, sample: str, path: str) -> None:
        order = sorted(range(len(self.svs)), key=lambda i: self.svs[i].key)
        with open(path, "w") as h:
            for i in order:
                for rd in self.samples[sample][i]:
                    m2 = "None" if self.svs[i].is_sbnd else rd.mapq2
                    h.write(f"{self.svs[i].key}\t{rd.rid}\t{rd.mapq1}\t{m2}\t{codes_to_str(rd.codes)}\n")

    def sv_order(self) -> List[int]:
        return sorted(range(len(self.svs)), key=lambda i: self.svs[i].key)

    # ---- straight into the device batch layout
    def to_batch(self, sample: str, score_ratio_thres: float = 1.2, var_read_min_mapq: int = 0,
                 sv_indices: Optional[List[int]] = None):
        bb = BatchBuilder(score_ratio_thres=score_ratio_thres, var_read_min_mapq=var_read_min_mapq)
        L = self.config.validate_sequence_length
        idxs = self.sv_order() if sv_indices is None else sv_indices
        for i in idxs:
            sv = self.svs[i]
            if sv.is_sbnd:
                var1, ref1 = sbnd_templates(sv.key, self.fasta, L)
                bb.add_sv([var1, None, ref1, None], True, False)
            else:
                var1, var2, ref1, ref2, short = canonical_templates(sv.key, self.fasta, L)
                bb.add_sv([var1, var2, ref1 if short else None, ref2 if short else None], False, short)
            for rd in self.samples[sample][i]:
                bb.add_read(rd.codes, rd.mapq1, None if sv.is_sbnd else rd.mapq2)
        return bb.finalize()


def _make_reference(rng: np.random.Generator, cfg: SynthConfig) -> Dict[str, np.ndarray]:
    contigs = {}
    for c in range(cfg.n_contigs):
        seq = rng.integers(0, 4, size=cfg.contig_len, dtype=np.uint8)
        n_runs = int(cfg.contig_len * cfg.n_run_rate / 50)
        for s in rng.integers(0, cfg.contig_len - 100, size=n_runs):
            seq[s:s + int(rng.integers(10, 90))] = 4
        contigs[f"chr{c + 1}"] = seq
    return contigs


def mutate_concat(rng: np.random.Generator, codes: np.ndarray, starts: np.ndarray, cfg: SynthConfig):
    """i.i.d. substitution / insertion / deletion errors over a concatenation of reads.
    codes: all true read sequences back to back, starts: start offset of each read. Returns (codes, starts, lens)."""
    n = len(codes)
    u = rng.random(n, dtype=np.float32)
    dele = u < cfg.del_rate
    sub = (u >= cfg.del_rate) & (u < cfg.del_rate + cfg.sub_rate) & (codes < 4)
    out = codes.copy()
    k = int(sub.sum())
    if k:
        out[sub] = (out[sub] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
    ins = rng.random(n, dtype=np.float32) < cfg.ins_rate
    reps = (~dele).astype(np.int8) + ins.astype(np.int8)
    expanded = np.repeat(out, reps)
    last = np.cumsum(reps, dtype=np.int64) - 1          # index of the last copy of every source base
    k = int(ins.sum())
    if k:
        expanded[last[ins]] = rng.integers(0, 4, size=k, dtype=np.uint8)
    csum = np.concatenate(([0], np.cumsum(reps, dtype=np.int64)))
    ends = np.concatenate((starts[1:], [n]))
    new_starts = csum[starts]
    new_lens = csum[ends] - new_starts
    return expanded, new_starts, new_lens


def _read_lengths(rng, cfg: SynthConfig, k: int) -> np.ndarray:
    ln = rng.gamma(cfg.read_len_shape, cfg.read_len_mean / cfg.read_len_shape, size=k)
    return np.clip(ln, cfg.read_len_min, cfg.read_len_max).astype(np.int64)


def _mapq(rng, cfg: SynthConfig) -> int:
    return int(rng.integers(0, 21)) if rng.random() < cfg.low_mapq_fraction else 60


def _window_read(rng, hap: np.ndarray, anchor: int, length: int, margin: int = 150) -> Tuple[int, int]:
    """A read of `length` drawn from haplotype `hap` that covers position `anchor` with `margin` on both sides."""
    length = min(length, len(hap))
    lo = max(0, anchor + margin - length)
    hi = min(len(hap) - length, anchor - margin)
    if hi < lo:
        lo = hi = max(0, min(len(hap) - length, anchor - length // 2))
    s = int(rng.integers(lo, hi + 1))
    return s, s + length


def make_workload(cfg: SynthConfig, n_sv: Optional[int] = None, samples: Tuple[str, ...] = ("tumor", "control")) -> Workload:
    """Deterministic (cfg.seed) synthetic workload with n_sv candidates (default cfg.n_sv)."""
    rng = np.random.default_rng(cfg.seed)
    n_sv = cfg.n_sv if n_sv is None else n_sv
    ref = _make_reference(rng, cfg)
    fasta = CodeFasta(ref)
    names = list(ref.keys())
    flank = cfg.read_len_max + 1000
    L = cfg.validate_sequence_length
    kinds = [k for k, _ in cfg.sv_mix]
    probs = np.array([p for _, p in cfg.sv_mix], dtype=float)
    probs /= probs.sum()

    pattern: List[Tuple[str, bool]] = []
    if cfg.stratified and not cfg.sbnd:
        # 40 slots in the proportions of sv_mix, the short ones spread over the DEL / DUP slots, in a fixed shuffled order
        P = 40
        counts = np.floor(probs * P + 1e-9).astype(int)
        for j in np.argsort(-(probs * P - counts))[:P - int(counts.sum())]:
            counts[j] += 1
        slots = [k for k, c in zip(kinds, counts) for _ in range(int(c))]
        dd = [j for j, k in enumerate(slots) if k in ("DEL", "DUP")]
        n_short = int(round(cfg.short_fraction * len(dd)))
        short_at = set(dd[int(round(q * len(dd) / max(n_short, 1)))] for q in range(n_short)) if dd else set()
        order = np.random.default_rng(cfg.seed ^ 0x5eed).permutation(P)
        pattern = [(slots[j], j in short_at) for j in order]
    svs: List[SynthSV] = []
    haps: List[Tuple[np.ndarray, int]] = []            # variant haplotype and junction position per SV
    bps: List[Tuple[Tuple[str, int], Optional[Tuple[str, int]]]] = []
    for i in range(n_sv):
        c1 = names[int(rng.integers(0, len(names)))]
        p1 = int(rng.integers(flank + 10, cfg.contig_len - 2 * flank - 10))
        if cfg.sbnd:
            direction = "+" if rng.random() < 0.5 else "-"
            clen = int(rng.integers(*cfg.sbnd_contig_range))
            tail = int(rng.integers(*cfg.sbnd_tail_range))
            novel = rng.integers(0, 4, size=max(clen, tail), dtype=np.uint8)
            contig = codes_to_str(novel[:clen])
            key = f"{c1},{p1},{direction},{contig[:L]},b_{i}"
            if direction == "+":
                left = ref[c1][p1 - flank:p1]
            else:
                left = revcomp_codes(ref[c1][p1 - 1:p1 - 1 + flank])
            haps.append((np.concatenate((left, novel)), len(left)))
            svs.append(SynthSV(key, True, "SBND"))
            bps.append(((c1, p1), None))
            continue
        kind = kinds[int(rng.choice(len(kinds), p=probs))]
        is_short = None
        if pattern:
            kind, is_short = pattern[i % len(pattern)]
        ins = ""
        c2, d1, d2 = c1, "+", "-"
        if kind == "DEL":
            u = rng.random()
            span = int(rng.integers(60, 280)) if (u < cfg.short_fraction if is_short is None else is_short) else int(rng.integers(500, 50000))
            p2 = p1 + span
        elif kind == "INS":
            p2 = p1 + 1
            ilen = int(rng.integers(*cfg.ins_len_range))
            ins = codes_to_str(rng.integers(0, 4, size=ilen, dtype=np.uint8))
        elif kind == "DUP":
            u = rng.random()
            span = int(rng.integers(60, 280)) if (u < cfg.short_fraction if is_short is None else is_short) else int(rng.integers(500, 50000))
            p2 = p1 + span
            d1, d2 = "-", "+"
        elif kind == "INV":
            p2 = p1 + int(rng.integers(500, 50000))
            d1, d2 = ("+", "+") if rng.random() < 0.5 else ("-", "-")
        else:  # TRA
            c2 = names[int(rng.integers(0, len(names)))]
            p2 = int(rng.integers(flank + 10, cfg.contig_len - 2 * flank - 10))
            d1 = "+" if rng.random() < 0.5 else "-"
            d2 = "+" if rng.random() < 0.5 else "-"
        p2 = min(p2, cfg.contig_len - flank - 10)
        key = f"{c1},{p1},{d1},{c2},{p2},{d2},{ins if ins else '---'},{'i' if kind == 'INS' else 'r'}_{i}"
        # variant haplotype, assembled like the reference builds its variant template but with long flanks
        left = ref[c1][p1 - flank:p1] if d1 == "+" else revcomp_codes(ref[c1][p1 - 1:p1 - 1 + flank])
        mid = np.frombuffer(ins.encode(), dtype=np.uint8)
        mid = np.searchsorted(_ASCII[:4], mid).astype(np.uint8) if len(mid) else np.zeros(0, dtype=np.uint8)
        if d1 == "-":
            mid = revcomp_codes(mid)
        right = ref[c2][p2 - 1:p2 - 1 + flank] if d2 == "-" else revcomp_codes(ref[c2][p2 - flank:p2])
        haps.append((np.concatenate((left, mid, right)), len(left) + len(mid) // 2))
        svs.append(SynthSV(key, False, kind))
        bps.append(((c1, p1), (c2, p2)))

    wl = Workload(cfg, fasta, svs)
    for sample in samples:
        vaf = cfg.tumor_vaf if sample == "tumor" else cfg.control_vaf
        true_chunks: List[np.ndarray] = []
        meta: List[Tuple[int, str, Optional[int], Optional[int], bool]] = []   # sv, id, mapq1, mapq2, revstrand
        for i, sv in enumerate(svs):
            (c1, p1), bp2 = bps[i]
            hap, junc = haps[i]
            n_var = int(rng.poisson(cfg.coverage * vaf))
            n_ref1 = int(rng.poisson(cfg.coverage * (1.0 - vaf)))
            n_ref2 = int(rng.poisson(cfg.coverage * (1.0 - vaf))) if bp2 is not None else 0
            lens = _read_lengths(rng, cfg, n_var + n_ref1 + n_ref2)
            k = 0
            for j in range(n_var):
                s, e = _window_read(rng, hap, junc, int(lens[k])); k += 1
                true_chunks.append(hap[s:e])
                meta.append((i, f"{sample[0]}{i}_v{j}", _mapq(rng, cfg), _mapq(rng, cfg), rng.random() < 0.5))
            for j in range(n_ref1):
                s, e = _window_read(rng, ref[c1], p1, int(lens[k])); k += 1
                true_chunks.append(ref[c1][s:e])
                also2 = bp2 is not None and bp2[0] == c1 and s <= bp2[1] + 100 and e >= bp2[1] - 100
                meta.append((i, f"{sample[0]}{i}_a{j}", _mapq(rng, cfg), _mapq(rng, cfg) if also2 else None,
                             rng.random() < 0.5))
            for j in range(n_ref2):
                c2, p2 = bp2
                s, e = _window_read(rng, ref[c2], p2, int(lens[k])); k += 1
                true_chunks.append(ref[c2][s:e])
                also1 = c2 == c1 and s <= p1 + 100 and e >= p1 - 100
                meta.append((i, f"{sample[0]}{i}_b{j}", _mapq(rng, cfg) if also1 else None, _mapq(rng, cfg),
                             rng.random() < 0.5))
        per_sv: List[List[SynthRead]] = [[] for _ in svs]
        if true_chunks:
            starts = np.zeros(len(true_chunks), dtype=np.int64)
            np.cumsum([len(c) for c in true_chunks[:-1]], out=starts[1:])
            noisy, nstarts, nlens = mutate_concat(rng, np.concatenate(true_chunks), starts, cfg)
            for (i, rid, m1, m2, rev), s0, ln in zip(meta, nstarts, nlens):
                codes = noisy[s0:s0 + ln]
                if rev:
                    codes = revcomp_codes(codes)
                if len(codes) == 0:
                    continue
                per_sv[i].append(SynthRead(rid, m1, m2, np.ascontiguousarray(codes)))
        wl.samples[sample] = per_sv
    return wl


def make_pair_sweep(n_pairs: int, m: int = 400, n_min: int = 1000, n_max: int = 50000, pool_reads: int = 2000,
                    seed: int = 20261019 + 5, match_fraction: float = 0.5, cfg: Optional[SynthConfig] = None):
    """Config 5: (read, template) pairs, template length m, read length uniform in [n_min, n_max]; half of the
    reads contain the template at ~8 % error, half are random. Pairs are drawn from a pool of reads.
    Returns (SeqPack-ready list of code arrays, tmpl_idx, read_idx) with templates first."""
    cfg = cfg or SynthConfig()
    rng = np.random.default_rng(seed)
    n_tmpl = max(1, min(pool_reads, 256))
    tmpls = [rng.integers(0, 4, size=m, dtype=np.uint8) for _ in range(n_tmpl)]
    chunks, starts, owner, pos = [], [], [], 0
    lens = rng.integers(n_min, n_max + 1, size=pool_reads)
    for r in range(pool_reads):
        t = int(rng.integers(0, n_tmpl))
        body = rng.integers(0, 4, size=int(lens[r]), dtype=np.uint8)
        if rng.random() < match_fraction and lens[r] > m + 10:
            at = int(rng.integers(0, lens[r] - m))
            src = tmpls[t] if rng.random() < 0.5 else revcomp_codes(tmpls[t])
            body[at:at + m] = src
        chunks.append(body); starts.append(pos); owner.append(t); pos += len(body)
    noisy, nstarts, nlens = mutate_concat(rng, np.concatenate(chunks), np.asarray(starts, dtype=np.int64), cfg)
    seqs = list(tmpls) + [np.ascontiguousarray(noisy[s:s + ln]) for s, ln in zip(nstarts, nlens)]
    read_pick = rng.integers(0, pool_reads, size=n_pairs)
    tmpl_idx = np.asarray([owner[r] for r in read_pick], dtype=np.uint32)
    read_idx = (read_pick + n_tmpl).astype(np.uint32)
    return seqs, tmpl_idx, read_idx
