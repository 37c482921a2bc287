#!/bin/bash
# One GPU-box visit: sanitizer on the small workload, the GPU test-suite, quick kernel timings. Everything under `timeout`.
mkdir -p gpurun_out
TAG=${1:-r2}
( timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py 2>&1 | tail -25 ) > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck: $(tail -2 gpurun_out/${TAG}_memcheck.log | tr '\n' ' ')"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc $(tail -1 gpurun_out/${TAG}_pytest.log)"
if [ $rc -ne 0 ]; then
  tail -40 gpurun_out/${TAG}_pytest.log
  for m in 1 2 4 6; do echo "== FB200_DEBUG_PATH=$m"; FB200_DEBUG_PATH=$m timeout 600 python -m pytest tests/test_gpu_encode.py -x -q 2>&1 | tail -3; done
fi
for W in cfg2 cfg2_l8 cfg3; do echo -n "$W: "; timeout 300 python bench.py --workload $W --kernels-only --steps 10 --warmup 3 2>&1 | tail -1; done | tee gpurun_out/${TAG}_quick.txt
