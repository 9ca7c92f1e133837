#!/bin/bash
# measurement aid: the bench line at N GPUs of the box (default 8), launched as the driver launches it
#   gpurun --gpus 8 -- 'bash tools/run_n8_check.sh 8'   ->  gpurun_out/r02_bench_n8.json
N=${1:-8}
if [ "$N" = 1 ]; then
  python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
fi
tail -c 300 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
for l in open("gpurun_out/r02_bench_n$N.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["traffic"], d["e2e"]["value"], d["extra"].get("reduce_ok"), d["scaling"])
        for k,v in d["extra"]["also"].items(): print(k, v["value"], v["ms_per_pass"], v["kernel_ms"], v.get("reduce_ok"))
PY
