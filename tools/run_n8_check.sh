#!/bin/bash
# measurement aid: the bench line at N GPUs of the box (default 8), as the driver launches it
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/r02c_bench_n$N.err
tail -c 300 gpurun_out/r02c_bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/r02c_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["extra"].get("reduce_ok"), d["scaling"])
        for k,v in d["extra"]["also"].items(): print(k, v["value"], v["ms_per_pass"], v["kernel_ms"], v.get("reduce_ok"))
PY
