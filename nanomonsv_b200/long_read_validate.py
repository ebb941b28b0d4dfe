"""GPU-backed compatibility layer for the reference's legacy module nanomonsv/long_read_validate.py.

The legacy module is imported by nothing in the reference (run.py:3-19) but is named by the north star, and it is where
the reference already has an engine switch: `ssw_check(query, target, use_ssw_lib)` (long_read_validate.py:12-17)
choosing between libssw (`ssw_check_ssw_lib`, :20-123) and parasail (`ssw_check_parasail`, :126-153). Here both names
resolve to the CUDA engine. The arithmetic is identical to the live path; the HOST logic differs and is reproduced
as it is written:

  * key = chr1,pos1,dir1,chr2,pos2,dir2,len(inseq)                     (:180, :245)
  * templates are NOT upper-cased                                      (:251-271)  -> soft-masked (lower-case) reference
    bases reach the aligner; my_seq.reverse_complement leaves them uncomplemented (explicit RC sequences are sent)
  * is_short_del_dup: same chromosome, (+,-) or (-,+), pos2 - pos1 + len(str(len(inseq))) < 100; the `== ''` on :160
    is a no-op comparison                                              (:158-166)
  * one combined predicate; in the short branch the var2 alternative tests ref1 twice and never ref2   (:316-326)
  * mapq filter requires the read to be known for the key            (:337-341)
  * defaults 1.4 / 0.2 / 0.8 / 10; long_read_validate_main passes 1.2 / 0.1 / 0.9 / 20   (:156, :423)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from . import _lib
from .count_sread_by_alignment import _aln_list, read_fasta
from .count_sread_by_alignment import ssw_check as _ssw_check_gpu
from .engine import Engine, SeqPack, default_engine
from .my_seq import needs_explicit_revcomp, reverse_complement


def ssw_check_parasail(query, target, engine: Optional[Engine] = None):
    """long_read_validate.py:126-153 -- same return value, computed on the GPU."""
    return _ssw_check_gpu(query, target, False, engine=engine)


def ssw_check_ssw_lib(query, target, engine: Optional[Engine] = None):
    """long_read_validate.py:20-123, the libssw route (dead code in the reference: its CLI flag is commented out,
    arg_parser.py:124-125). The host side of that route -- strand rule, coordinate conversion (:102-113) -- is the same as
    the parasail route's; that libssw's ssw_align(flag=1, maskLen=m/2) returns the same cells as parasail.ssw is NOT
    verified here (libssw is not in this image): this name runs the same CUDA engine as ssw_check_parasail."""
    return _ssw_check_gpu(query, target, True, engine=engine)


def ssw_check(query, target, use_ssw_lib, engine: Optional[Engine] = None):
    """long_read_validate.py:12-17: the engine switch of the reference; both branches are the CUDA engine here."""
    return ssw_check_ssw_lib(query, target, engine) if use_ssw_lib else ssw_check_parasail(query, target, engine)


def legacy_key(F: Sequence[str]) -> str:
    inseq = "" if F[6] == "---" else F[6]
    return ",".join([F[0], str(int(F[1])), F[2], F[3], str(int(F[4])), F[5], str(len(inseq))])


def is_short_del_dup(key: str) -> bool:
    k = key.split(",")
    same = k[0] == k[3]
    span = int(k[4]) - int(k[1]) + len(k[6])        # len() of the decimal string, as written in the reference
    return same and ((k[2] == "+" and k[5] == "-") or (k[2] == "-" and k[5] == "+")) and span < 100


def legacy_templates(F: Sequence[str], fasta, L: int = 200) -> List[str]:
    """[variant_seq_1, variant_seq_2, reference_local_seq_1, reference_local_seq_2] (:241-281), no upper-casing."""
    chr1, pos1, dir1, chr2, pos2, dir2 = F[0], int(F[1]), F[2], F[3], int(F[4]), F[5]
    inseq = "" if F[6] == "---" else F[6]
    ref1 = fasta.fetch(chr1, max(pos1 - L - 1, 0), pos1 + L - 1)
    ref2 = fasta.fetch(chr2, max(pos2 - L - 1, 0), pos2 + L - 1)
    if dir1 == "+":
        variant = fasta.fetch(chr1, max(pos1 - L - 1, 0), pos1 - 1) + inseq
    else:
        variant = reverse_complement(fasta.fetch(chr1, pos1 - 1, pos1 + L - 1)) + reverse_complement(inseq)
    if dir2 == "-":
        variant += fasta.fetch(chr2, pos2 - 1, pos2 + L - 1)
    else:
        variant += reverse_complement(fasta.fetch(chr2, max(pos2 - L - 1, 0), pos2 - 1))
    n = min(2 * L, len(variant))
    return [variant[:n], variant[len(variant) - n:], ref1, ref2]


def long_read_validate_from_reads(sv_rows: Sequence[Sequence[str]], key2reads: Dict[str, List[Tuple[str, str]]],
                                  key2rname2mapq: Dict[str, Dict[str, list]], fasta, variant_sread_min_mapq=0,
                                  validate_sequence_length=200, score_ratio_thres=1.4, start_pos_thres=0.2,
                                  end_pos_thres=0.8, var_ref_margin_thres=10, engine: Optional[Engine] = None):
    """The part of long_read_validate_by_alignment after read gathering (:234-417): returns
    [key2sread_count, key2sread_count_all, key2sread_info]. key2reads[key] = [(qname, sequence)] in file order."""
    eng = engine or default_engine()
    key2tmpl: Dict[str, List[str]] = {}
    for F in sv_rows:
        key2tmpl[legacy_key(F)] = legacy_templates(F, fasta, validate_sequence_length)

    # one device batch: every (template, read) pair of every key that has reads
    pack = SeqPack()
    t_idx, rc_idx, r_idx, owner = [], [], [], []
    for key, reads in key2reads.items():
        if key not in key2tmpl or not reads:
            continue
        dedup: Dict[str, str] = {}
        for rid, seq in reads:          # the FASTA round trip of the reference: last record of an id wins
            if seq:
                dedup[rid.split()[0]] = seq
        ridx = {rid: pack.add(seq) for rid, seq in dedup.items()}
        n_t = 4 if is_short_del_dup(key) else 2
        for s in range(n_t):
            tmpl = key2tmpl[key][s]
            ti = pack.add(tmpl)
            rci = pack.add(reverse_complement(tmpl)) if needs_explicit_revcomp(tmpl) else _lib.NMSV_NONE
            for rid, ri in ridx.items():
                t_idx.append(ti); rc_idx.append(rci); r_idx.append(ri); owner.append((key, s, rid))
    res = eng.ssw_check_batch(pack, t_idx, r_idx, rc_idx) if t_idx else []
    info: Dict[str, List[Dict[str, list]]] = {}
    for (key, s, rid), a in zip(owner, res):
        info.setdefault(key, [{}, {}, {}, {}])[s][rid] = _aln_list(a)

    key2count, key2count_all, key2info = {}, {}, {}
    t = var_ref_margin_thres
    for key, per in info.items():
        v1, v2, r1, r2 = per
        len1, len2 = len(key2tmpl[key][0]), len(key2tmpl[key][1])
        names = list(dict.fromkeys(list(v1.keys()) + list(v2.keys())))

        def cover(a, ln):
            return a[0] > score_ratio_thres * ln and a[1] < start_pos_thres * ln and a[2] > end_pos_thres * ln

        if is_short_del_dup(key):
            supp = [n for n in names if
                    (cover(v1[n], len1) and v1[n][0] >= r1[n][0] + t and v1[n][0] >= r2[n][0] + t) or
                    (cover(v2[n], len2) and v2[n][0] >= r1[n][0] + t and v2[n][0] >= r1[n][0] + t)]   # ref1 twice, as written
        else:
            supp = [n for n in names if cover(v1[n], len1) or cover(v2[n], len2)]
        mq = key2rname2mapq.get(key, {})
        supp = [n for n in supp if n in mq and mq[n][0] is not None and mq[n][0] >= variant_sread_min_mapq
                and mq[n][1] is not None and mq[n][1] >= variant_sread_min_mapq]
        key2count[key] = len(supp)
        key2count_all[key] = len(names)
        key2info[key] = {n: v1[n] + v2[n] for n in supp}
    return [key2count, key2count_all, key2info]


def long_read_validate_by_alignment(sv_file, output_file, bam_file, reference, use_ssw_lib, debug,
                                    variant_sread_min_mapq=0, validate_sequence_length=200, score_ratio_thres=1.4,
                                    start_pos_thres=0.2, end_pos_thres=0.8, var_ref_margin_thres=10,
                                    engine: Optional[Engine] = None):
    """Same signature as the reference (:156). Read gathering needs pysam (lazy import), alignment runs on the GPU."""
    import os
    import pysam

    bam = pysam.AlignmentFile(bam_file, "rb")
    sv_rows, rname2key, key2rname2mapq = [], {}, {}
    with open(sv_file) as hin:
        for line in hin:
            if line.startswith("#") or line.startswith("Chr_1"):
                continue
            F = line.rstrip("\n").split("\t")
            sv_rows.append(F)
            key = legacy_key(F)
            per = key2rname2mapq.setdefault(key, {})
            for which, (c, p) in enumerate(((F[0], int(F[1])), (F[3], int(F[4])))):
                for read in bam.fetch(c, max(p - 100, 0), p + 100):
                    rname2key.setdefault(read.qname, []).append(key)
                    per.setdefault(read.qname, [None, None])[which] = read.mapping_quality
    key2reads: Dict[str, List[Tuple[str, str]]] = {}
    for read in bam.fetch():
        if read.is_supplementary or read.is_secondary or read.is_duplicate or read.qname not in rname2key:
            continue
        seq = read.query_sequence
        if read.is_reverse:
            seq = reverse_complement(seq)
        for key in set(rname2key[read.qname]):
            key2reads.setdefault(key, []).append((read.qname, seq))
    bam.close()
    fasta = pysam.FastaFile(os.path.abspath(reference))
    return long_read_validate_from_reads(sv_rows, key2reads, key2rname2mapq, fasta, variant_sread_min_mapq,
                                         validate_sequence_length, score_ratio_thres, start_pos_thres, end_pos_thres,
                                         var_ref_margin_thres, engine)


def long_read_validate_main(result_file, tumor_bam, output, sread_file, reference, control_bam, var_read_min_mapq,
                            use_ssw_lib, debug, engine: Optional[Engine] = None):
    """long_read_validate.py:421-463 -- writes F[:7] + counts and the sread file, formats unchanged."""
    kw = dict(variant_sread_min_mapq=var_read_min_mapq, score_ratio_thres=1.2, start_pos_thres=0.1, end_pos_thres=0.9,
              var_ref_margin_thres=20, engine=engine)
    cnt_t, all_t, info_t = long_read_validate_by_alignment(result_file, output, tumor_bam, reference, use_ssw_lib, debug, **kw)
    if control_bam is not None:
        cnt_c, all_c, _ = long_read_validate_by_alignment(result_file, output, control_bam, reference, use_ssw_lib, debug, **kw)
    with open(result_file) as hin, open(output, "w") as hout:
        for line in hin:
            if line.startswith("#") or line.startswith("Chr_1"):
                continue
            F = line.rstrip("\n").split("\t")
            key = legacy_key(F)
            cols = F[:7] + [str(all_t.get(key, 0)), str(cnt_t.get(key, 0))]
            if control_bam is not None:
                cols += [str(all_c.get(key, 0)), str(cnt_c.get(key, 0))]
            print("\t".join(cols), file=hout)
    with open(sread_file, "w") as hout:
        for key, per in info_t.items():
            for sread, vals in per.items():
                print(key + "\t" + sread + "\t" + "\t".join(str(x) for x in vals), file=hout)
