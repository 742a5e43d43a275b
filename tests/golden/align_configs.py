"""Parameter sets of `mmseqs align` used by make_align_golden.py (reference side) and tests/test_alignment_batch.py (device side)."""
# name -> keyword arguments of Ref.align_query (the parameters of `mmseqs align` that reach the per-query loop)
CONFIGS = {
    "default_a": dict(sw_mode=2, eval_thr=1e-3),                                          # easy-search default (-a)
    "mode1_cov": dict(sw_mode=1, eval_thr=10.0, cov_thr=0.5, cov_mode=0, add_backtrace=False),
    "mode0": dict(sw_mode=0, eval_thr=1e3, add_backtrace=False),
    "strict": dict(sw_mode=2, eval_thr=1e-5, cov_thr=0.8, cov_mode=2, seq_id_thr=0.3, seq_id_mode=1, aln_len_thr=30,
                   max_accept=5, max_reject=3),
    "target_cov_nobias": dict(sw_mode=2, eval_thr=1.0, cov_thr=0.6, cov_mode=1, comp_bias=False, seq_id_mode=2, compress=False),
    "len_modes": dict(sw_mode=1, eval_thr=1e-2, cov_thr=0.7, cov_mode=5, add_backtrace=False, max_reject=10),
}

# nucleotide searches (Ref.align_query_nucl / b200_align_batch_nucl); gap 5/2, zdrop 40 as in the reference's defaults
NUCL_CONFIGS = {
    "nucl_default": dict(eval_thr=1e-3),
    "nucl_strict": dict(eval_thr=1e-10, cov_thr=0.8, cov_mode=2, seq_id_thr=0.9, seq_id_mode=1, aln_len_thr=50, max_accept=2, max_reject=2),
    "nucl_loose": dict(eval_thr=1e3, cov_thr=0.3, cov_mode=1, compress=False),
}
