# Round 2, first GPU call: validate what round 1 left pending, try compute-sanitizer, take a baseline bench.
set -o pipefail
mkdir -p gpurun_out
BZ_VALIDATE_PENDING=1 timeout 900 python -m pytest tests/test_gpu_prover.py tests/test_gpu_arith.py -x -q -k "narrow_geometry or binary_gcd" 2>&1 | tail -5 | tee gpurun_out/r2_pending.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitizer.log 2>&1; echo "sanitizer rc=$?" | tee -a gpurun_out/r2_sanitizer.log
tail -15 gpurun_out/r2_sanitizer.log
timeout 900 python bench.py > gpurun_out/r2_bench0.log 2>&1; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_bench0.log
