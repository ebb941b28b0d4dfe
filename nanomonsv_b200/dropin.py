"""Switches an installed reference package (nanomonsv v0.8.0) over to the GPU paths of this repository without editing it:

    python -m nanomonsv_b200.dropin get  <the reference's own arguments>
    python -m nanomonsv_b200.dropin validate ...

or, from Python, `nanomonsv_b200.dropin.install()` before the reference's functions are called. What is replaced:

  reference symbol (file:line)                                              replaced by
  count_sread_by_alignment            count_sread_by_alignment.py:401       nanomonsv_b200.count_sread_by_alignment.count_sread_by_alignment
  locate_bp                           locate_bp.py:88                       nanomonsv_b200.locate_bp.locate_bp (one sw_jump batch per file)
  locate_bp_sbnd                      locate_bp_sbnd.py:43                  nanomonsv_b200.locate_bp_sbnd.locate_bp_sbnd
  Consensus_generator.generate_paf_file   generate_consensus.py:100         nanomonsv_b200.generate_consensus.generate_paf_file
  Sv_filterer.filter_sv_insertion_match   post_proc.py:84                   verdicts of nanomonsv_b200.sv_filter, all pairs of the list in one batch
  smith_waterman.sw_jump              smith_waterman.py:5                   nanomonsv_b200.smith_waterman.sw_jump
  the `parasail` module               imported by five reference modules    nanomonsv_b200.ssw (matrix_create / ssw), only if parasail is absent

`run.py` binds the stage functions with `from .module import *`, so the names are replaced in `nanomonsv.run` as well as in their
home modules. The reference's loops, filters, file formats and command line stay the reference's own code."""
from __future__ import annotations

import sys


def install(engine=None) -> dict:
    """Returns {patched name: replacement}. Needs the reference package `nanomonsv` on sys.path (and pysam for real BAM / FASTA
    files; parasail is not needed any more)."""
    from . import count_sread_by_alignment as g_count, generate_consensus as g_gc, locate_bp as g_lb, locate_bp_sbnd as g_lbs
    from . import smith_waterman as g_sw, ssw as g_ssw, sv_filter as g_filter
    if "parasail" not in sys.modules:
        try:
            import parasail  # noqa: F401  (a real one stays in place for whatever this repository does not replace)
        except ImportError:
            sys.modules["parasail"] = g_ssw
    import nanomonsv.count_sread_by_alignment as r_count
    import nanomonsv.generate_consensus as r_gc
    import nanomonsv.locate_bp as r_lb
    import nanomonsv.locate_bp_sbnd as r_lbs
    import nanomonsv.post_proc as r_pp
    import nanomonsv.run as r_run
    import nanomonsv.smith_waterman as r_sw
    done = {}

    def put(holders, name, fn):
        for h in holders:
            setattr(h, name, fn)
        done[name] = fn

    put((r_count, r_run), "count_sread_by_alignment", g_count.count_sread_by_alignment)
    put((r_lb, r_run), "locate_bp", g_lb.locate_bp)
    put((r_lb,), "get_refined_bp", g_lb.get_refined_bp)
    put((r_lbs, r_run), "locate_bp_sbnd", g_lbs.locate_bp_sbnd)
    put((r_lbs,), "get_refined_bp_sbnd", g_lbs.get_refined_bp_sbnd)
    put((r_sw,), "sw_jump", g_sw.sw_jump)

    def generate_paf_file(self, query_fasta, target_fasta, output_file):
        return g_gc.generate_paf_file(query_fasta, target_fasta, output_file, self.parasail_error)
    put((r_gc.Consensus_generator,), "generate_paf_file", generate_paf_file)

    def filter_sv_insertion_match(self, sv, ins, filter_item="Duplicate_with_insertion"):
        # the reference's loop (apply_filters :226-232) keeps deciding which pairs are looked at and in which order; what a pair
        # decides depends on its two sequences only, so the verdicts of all pairs of the list are worked out on first use, as
        # one batch of alignments
        cache = getattr(self, "_nmsv_verdicts", None)
        if cache is None or cache[0] != len(self.sv_list):
            examined, pairs = [], []
            for a, b in g_filter._ordered_pairs(self.sv_list, self.min_indel_size):
                segs = g_filter.insertion_match_segments(a, b, self.reference_h, self.min_indel_size, self.bp_dist_margin,
                                                         self.validate_seg_len)
                if segs is not None:
                    examined.append((id(a), id(b), len(segs[0])))
                    pairs += [(segs[0], segs[1]), (g_filter.reverse_complement(segs[0]), segs[1])]
            res = g_ssw.ssw_batch(pairs, 3, 1, g_ssw.matrix_create("ACGT", 2, -2), engine=engine) if pairs else []
            table = {(ia, ib): g_filter.insertion_match_verdict(res[2 * k], res[2 * k + 1], n) for k, (ia, ib, n) in enumerate(examined)}
            cache = (len(self.sv_list), table)
            self._nmsv_verdicts = cache
        if cache[1].get((id(sv), id(ins)), False):
            sv.filter.append(filter_item)
    put((r_pp.Sv_filterer,), "filter_sv_insertion_match", filter_sv_insertion_match)
    return done


def main() -> None:
    install()
    import nanomonsv
    nanomonsv.main()


if __name__ == "__main__":
    main()
