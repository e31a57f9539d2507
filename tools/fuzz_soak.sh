#!/bin/bash
# Soak run of the randomised GPU parity sweeps: tests/test_gpu_fuzz.py once per seed offset (HCJ_FUZZ_SEED).
#   tools/fuzz_soak.sh FIRST LAST [LOG]      e.g. under gpurun: tools/fuzz_soak.sh 1 30 gpurun_out/fuzz_soak.log
first=${1:-1}; last=${2:-10}; log=${3:-/dev/stdout}
fail=0
for s in $(seq "$first" "$last"); do
  if HCJ_FUZZ_SEED=$s timeout 300 python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q > /tmp/fuzz_$s.log 2>&1; then
    echo "seed $s: $(tail -1 /tmp/fuzz_$s.log)" >> "$log"
  else
    fail=$((fail + 1)); echo "seed $s: FAILED" >> "$log"; tail -40 /tmp/fuzz_$s.log >> "$log"
  fi
done
echo "failures: $fail of $((last - first + 1)) seeds" >> "$log"
exit $((fail > 0))
