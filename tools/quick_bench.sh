#!/bin/bash
# quick device-resident timing of the standard workloads (kernels only); prints one JSON line each
for W in ${@:-cfg2 cfg2_l8 cfg3}; do echo -n "$W: "; python bench.py --workload $W --kernels-only --steps 10 --warmup 3 2>&1 | tail -1; done
