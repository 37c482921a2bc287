#!/bin/bash
# ncu --set full of the hot kernels (last of 4 launch sets) for the given workloads + the default bench line.
TAG=${1:-r2}; shift
mkdir -p gpurun_out
for W in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_}" --launch-skip ${SKIP:-15} --launch-count ${COUNT:-5} \
     -f -o gpurun_out/${TAG}_${W} python bench.py --workload $W --kernels-only --steps 1 --warmup 3 > gpurun_out/${TAG}_${W}_ncu.log 2>&1
  echo "ncu $W rc=$?"
done
