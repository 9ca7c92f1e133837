#!/bin/bash
# measurement aid: everything profiles/r02_* is made from, in ONE gpurun call on one GPU (tools/summarize_r02.py turns the captures
# into the committed summaries afterwards).  Each ncu pass runs only after the same command has exited 0 without ncu.
#   gpurun --timeout 1800 -- 'bash tools/profile_round.sh'
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || exit 1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for wl in c3 c5; do
  WORKLOAD=$wl python tools/profile_lightcone.py > gpurun_out/r02_plain_$wl.log 2>&1 || exit 1
  WORKLOAD=$wl ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_$wl.csv \
    python tools/profile_lightcone.py > gpurun_out/r02_ncu_$wl.log 2>&1 || exit 1
done
PGROUPS=8 python tools/profile_lightcone.py > gpurun_out/r02_plainfull_c3.log 2>&1 || exit 1
PGROUPS=8 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"deposit_pipelined|bin_histogram|bin_scatter|tile_deposit" -c 4 -f -o gpurun_out/r02_full_c3 \
  python tools/profile_lightcone.py > gpurun_out/r02_ncufull_c3.log 2>&1 || exit 1
# the full-set run rewrites the partial meta only; keep the light-cone metas
cuobjdump -sass slicer_b200/_build/libslicer_b200.so > gpurun_out/r02_sass_full.txt 2>/dev/null
echo done
