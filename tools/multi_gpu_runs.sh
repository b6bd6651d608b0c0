#!/bin/bash
# Multi-GPU records of one box: tools/multi_gpu_runs.sh N tag  (run under `gpurun --gpus N`); writes gpurun_out/<tag>_*.json
N=$1; tag=$2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline "$@"; }
run > gpurun_out/${tag}_weak_${N}gpu.json 2> gpurun_out/${tag}_weak_${N}gpu.err
run --scaling strong --scans 200 > gpurun_out/${tag}_strong200_${N}gpu.json 2> gpurun_out/${tag}_strong200_${N}gpu.err
if [ "$3" == "sweep" ]; then
  for pts in 1000 20000; do
    run --points $pts > gpurun_out/${tag}_weak_${pts}pts_${N}gpu.json 2> gpurun_out/${tag}_weak_${pts}pts_${N}gpu.err
  done
fi
for f in gpurun_out/${tag}_*_${N}gpu*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], d["n_gpus"], "gpus", round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]), "%.1f ms" % d["ms_per_step"], d["scaling"], d["config"]["workload"][:60])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
