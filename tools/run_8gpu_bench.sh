#!/bin/bash
# the default workload on 8 GPUs (read-pair sharding, weak scaling) with its config-4 block (5,000 x 4 Mbp genomes range-partitioned)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2n_bench_default_n8.json 2> gpurun_out/r2n_bench_default_n8.err
tail -c 800 gpurun_out/r2n_bench_default_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench_default_n8.json').read().strip().splitlines()[-1])
print("N=8 value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"]); print(json.dumps(d.get("configs"))[:3000])
PY
