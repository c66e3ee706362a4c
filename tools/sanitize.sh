#!/bin/bash
# compute-sanitizer memcheck + racecheck over a curated set of GPU tests that covers every ring kernel
# (TMA Cholesky ring, generic chain sweep fwd/bwd, parallel-in-time paths, in-place / misaligned arrays,
# Kalman sweeps, large-block kernels incl. the half-warp parallel-in-time path, adjoint sweeps, in-kernel
# sampling, peer exchange, the tensor-map sweep engine, the warp-per-chain kernels for D = 9..32).  Usage (on a GPU box): tools/sanitize.sh [outdir]
# racecheck runs twice: with the default completion of the ragged element copies
# (cp.async.mbarrier.arrive.noinc -- the tool does not model that path and flags the consumers' reads) and with
# tuning knob 12 = 1 (the issuing thread waits for its copies and arrives itself: same results, tool-visible).
out=${1:-gpurun_out}
mkdir -p "$out"
SEL='test_cholesky_reads_lower_triangle_only_and_in_place_alias or test_cholesky_parallel_in_time_segments_alias_and_failure or test_config2_shape_slice or test_solve_vs_oracle_tight or test_inverse_subset_with_subdiag_vs_oracle_tight or test_upper_diagonal_lower_vs_oracle_tight or test_in_place_factorisation_and_failure_report or test_config4_sum_kernel or test_log_likelihood_matches_reference_kalman_filter or test_mid_size_batch or test_cvi_style_site_update or test_naturals_to_ssm_params_parallel_in_time_short_segments or test_dense_gp_closed_form or test_marginals_parallel_in_time_segment_length_knob or test_sparse_sites or test_cholesky_solve_logdet_vs_oracle or (test_parallel_in_time_large_blocks and 17 and float64) or test_marginals_gradients_match or test_cholesky_solve_logdet_gradients or test_in_kernel_sampling or (test_time_sharded_exchange and 3) or (test_naturals_to_ssm_params_tensor_map_engine and 2-dtype0) or (test_forward_moment_sweeps_tensor_map_engine and 2-dtype0) or test_tensor_map_engine_declines or (test_transforms_match_oracle_above_eight_dimensions and 17) or (test_block_operators_above_eight_dimensions and 17) or (test_kalman_log_likelihood_above_eight_dimensions and 17) or test_naturals_to_ssm_params_reverse_mode or test_sample_reverse_mode'
run() {  # name, tool, env
  env $3 timeout 1500 compute-sanitizer --tool $2 --print-limit 20 --error-exitcode 0 \
    python -m pytest tests -m gpu -q -x -k "$SEL" -p no:cacheprovider > "$out/sanitizer_$1.log" 2>&1
  echo "rc=$?" >> "$out/sanitizer_$1.log"
  echo "== $1"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" "$out/sanitizer_$1.log" | tail -4
}
run memcheck memcheck "MF_X=0"
run racecheck_default racecheck "MF_X=0"
run racecheck_knob12 racecheck "MF_TUNING=12=1"
