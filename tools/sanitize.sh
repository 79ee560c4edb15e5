#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over a reduced -m gpu selection: K1 (uint8 + float32 / TMA staging, all three
# pass forms), the Hamming evaluator (stage A / S / B with and without the stash, 1-4 threads per query, the select
# pipeline incl. its retry round and the pool-overflow fallback), the tensor-core k-NN scorer with the fused selection,
# the fused resize kernel.  Run under gpurun; logs land in gpurun_out/sanitize_*.log (summaries are copied to profiles/).
set -u
OUT=gpurun_out
mkdir -p $OUT
run() {   # tool tag pytest-args...
    local tool=$1 tag=$2; shift 2
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 86 --print-limit 20 python -m pytest -x -q "$@" > $OUT/sanitize_${tool}_${tag}.log 2>&1
    echo "$tool $tag: exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitize_${tool}_${tag}.log | tail -1) : $(grep -E ' passed| failed| error' $OUT/sanitize_${tool}_${tag}.log | tail -1)"
}
for tool in memcheck racecheck; do
    run $tool select tests/test_gpu_select.py -k "select_pipeline_matches_exact_oracle and (37-2000 or 33-4096 or 40-5000 or 3-777) or misleading or class_sorted"
    run $tool eval tests/test_gpu_eval.py -k "maphashing_and_ranking_match_exact_oracle and (37-500-64-24-50 or 20-3000-128 or 40-5000-48 or 7-1029) or stage_a_threads_per_query_agree or host_buffer_entry_point_directly"
    run $tool engine tests/test_gpu_engine.py -k "engine_matches_exact_oracle and (64-5001 or 300-40000) or packed_host_entry_point"
    run $tool swt tests/test_gpu_swt.py -k "pass_forms or uint8_staging or constant_and_linearity or host_buffer or (matches_oracle and (256 or 224-224-haar or db7 or coif1))"
    run $tool knn tests/test_gpu_eval.py -k "knn_tensor_core_scorer_matches_simt_scorer or knn_lists_longer and 4097 or sharded_knn_merge and 500"
    run $tool resize tests/test_gpu_resize.py
done
