#!/bin/bash
# compute-sanitizer over the non-tensor kernels (LDATI, event frames, pre-processing) on the small test shapes
# ("race detection / sanitizers: none upstream").  memcheck: out-of-bounds / misaligned accesses;
# racecheck: shared-memory hazards (the one-sweep sort's counter / stage aliasing, the word transpose of the record
# writer); synccheck: barrier misuse.  The tcgen05 conv kernels are exercised by memcheck only through the small
# forward of the pipeline test.  Slow (10-50x): run on demand,
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh gpurun_out/sanitize'
set -u
OUT=${1:-gpurun_out/sanitize}
mkdir -p "$OUT"
SEL='events_bit_exact_vs_oracle or other_frame_rates or injected_draws or empty_and_negative or output_buffer_of_any_alignment'
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 --log-file "$OUT/ldati_$tool.log" \
      python -m pytest tests/test_gpu_ldati.py tests/test_gpu_event_frames.py -m gpu -q -x -k "$SEL or stages_bit_exact or count_pass_writes" \
      > "$OUT/ldati_$tool.pytest.txt" 2>&1
  echo "$tool: exit $? -- $(grep -c 'ERROR SUMMARY' "$OUT/ldati_$tool.log") summary line(s): $(grep 'ERROR SUMMARY' "$OUT/ldati_$tool.log" | tail -1)"
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file "$OUT/pipeline_memcheck.log" \
    python -m pytest tests/test_gpu_pipeline.py -m gpu -q -x -k "batch_runner_overlapped" > "$OUT/pipeline_memcheck.pytest.txt" 2>&1
echo "pipeline memcheck: exit $? $(grep 'ERROR SUMMARY' "$OUT/pipeline_memcheck.log" | tail -1)"
