# compute-sanitizer over the small GPU parity tests on the final kernels (racecheck: shared-memory hazards; memcheck:
# out-of-bounds / misaligned accesses; synccheck: divergent barriers).  Output summarised into profiles/r2_sanitizer.txt
SEL='test_predict_topn_matches_oracle or test_predict_lists_only or test_fit_matches_canonical_oracle_on_golden_inputs or test_fit_total_ties or test_fit_items_seen_by_more_than_65535_users or test_predict_heavy_user_limb_chunks or test_metric_unit_vectors or test_top_k_ranks'
for tool in racecheck memcheck synccheck; do
  echo "==== compute-sanitizer --tool $tool"
  timeout 2400 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "exit code $?"
  grep -E "passed|failed|error" gpurun_out/sanitizer_$tool.log | tail -3
  grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard" gpurun_out/sanitizer_$tool.log | sort | uniq -c | head -10
done
