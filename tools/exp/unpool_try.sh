#!/bin/bash
# quick matrix of unpool_fused configurations (debugging aid)
cd /root/repo
run() { echo "== P=$P N=$N $*"; env "$@" timeout 60 python tools/unpool_bench.py --fused-only --iters 20 --pairs $P --points $N 2>&1 | grep "unpool_fused:\|ktime\|pair-res\|kernel alone" | tail -3; }
P=149 N=2000 run LMPCR_UNPOOL_DEBUG=1
P=296 N=1024 run LMPCR_UNPOOL_DEBUG=1
P=296 N=2000 run LMPCR_UNPOOL_DEBUG=1
P=296 N=2000 run LMPCR_UNPOOL_PARTS=1
P=296 N=2000 run LMPCR_UNPOOL_PARTS=2
P=296 N=5000 run A=1
P=100 N=2000 run A=1
