#!/bin/bash
# two-GPU checks of the partitioned path (gpurun --gpus 2): the default workload with its config-4 block
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2k_bench_default_n2.json 2> gpurun_out/r2k_bench_default_n2.err; tail -c 300 gpurun_out/r2k_bench_default_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k_bench_default_n2.json').read().strip().splitlines()[-1])
print("N=2 value", d["value"], "e2e", d["e2e"]["value"]); print(json.dumps(d.get("configs"))[:2500])
PY
