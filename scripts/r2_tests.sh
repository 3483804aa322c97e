# full GPU parity suite with durations
set -o pipefail
mkdir -p gpurun_out
BZ_VALIDATE_PENDING=1 timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 | tee gpurun_out/r2_tests.log
