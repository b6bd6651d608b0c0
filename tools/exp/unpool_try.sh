#!/bin/bash
# quick matrix of unpool_fused configurations (timing experiments)
cd /root/repo
run() { echo "== P=$P N=$N $*"; env "$@" timeout 60 python tools/unpool_bench.py --fused-only --iters 10 --pairs $P --points $N 2>&1 | grep "unpool_fused:\|kernel alone" | tail -3; }
P=296 N=2000 run LMPCR_UNPOOL_EPI_SLEEP=0
P=296 N=2000 run LMPCR_UNPOOL_EPI_SLEEP=1
P=296 N=2000 run LMPCR_UNPOOL_EPI_SLEEP=0
P=296 N=2000 run LMPCR_UNPOOL_EPI_SLEEP=1
