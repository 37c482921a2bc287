#!/bin/bash
# pipeline granularity sweep: device-resident step time vs number of sub-batches per call
for W in cfg2 cfg2_l8; do for C in 1 2 3 4 6; do echo -n "$W chunks=$C: "; FB200_BENCH_MAXBLOCKS=10000 FB200_PIPE_CHUNKS=$C python bench.py --workload $W --kernels-only --steps 10 --warmup 3 2>&1 | tail -1; done; done
