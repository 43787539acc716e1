#!/bin/bash
# compute-sanitizer over the hot path: memcheck, racecheck and synccheck on smoke() and on the two
# structural parity tests of the tile mixer (dense polyphony, many batches in one plan across the internal streams), the
# adversarial test of the branch-and-bound peak pass and the FX tests (FX streams, the lane = (row, stage) dynamics kernel).
# Logs go to gpurun_out/sanitizer_<tool>_<what>.txt; copy the summaries into profiles/.
#   usage: tools/gpu_sanitize.sh [tool ...]      (default: memcheck racecheck synccheck)
mkdir -p gpurun_out
TOOLS=${@:-memcheck racecheck synccheck}
TESTS="tests/test_gpu_parity.py::test_dense_polyphony_deterministic_and_correct tests/test_gpu_parity.py::test_batches_in_one_plan_equal_batch_by_batch tests/test_gpu_parity.py::test_peak_pass_on_banks_that_defeat_its_pruning tests/test_gpu_parity.py::test_mel_fast_and_generic_filterbank_paths tests/test_gpu_fx.py tests/test_gpu_projection.py tests/test_gpu_handles.py"
for tool in $TOOLS; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 0 \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_smoke.txt 2>&1
  echo "== $tool smoke: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_smoke.txt | tail -1)"
  timeout 1500 compute-sanitizer --tool $tool --print-limit 30 --error-exitcode 0 \
      python -m pytest -x -q -m gpu $TESTS > gpurun_out/sanitizer_${tool}_tests.txt 2>&1
  echo "== $tool tests: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_${tool}_tests.txt | tail -2 | tr '\n' ' ')"
done
