#!/bin/bash
# ncu --set full of the hot kernels for the given workloads. The .ncu-rep files are large (gpurun merges at most 64 MiB
# back), so the raw page and the per-kernel source pages are exported on the box as gzip'ed CSV and the report is dropped
# unless KEEP=1.   usage: [KREGEX=..] [SKIP=n] [COUNT=n] [KEEP=1] tools/gpu_prof.sh <tag> <workload>...
TAG=${1:-r2}; shift
mkdir -p gpurun_out
for W in "$@"; do
  REP=gpurun_out/${TAG}_${W}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_}" --launch-skip ${SKIP:-15} --launch-count ${COUNT:-5} \
     -f -o $REP python bench.py --workload $W --kernels-only --steps 1 --warmup 3 > ${REP}_ncu.log 2>&1
  echo "ncu $W rc=$?"
  ncu -i $REP.ncu-rep --page raw --csv 2>/dev/null | gzip > ${REP}_raw.csv.gz
  ncu -i $REP.ncu-rep --page source --csv 2>/dev/null | gzip > ${REP}_source.csv.gz
  ls -la $REP.ncu-rep ${REP}_raw.csv.gz ${REP}_source.csv.gz
  [ "${KEEP:-0}" = 1 ] || rm -f $REP.ncu-rep
done
