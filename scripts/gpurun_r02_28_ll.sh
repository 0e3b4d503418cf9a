#!/bin/bash
# launch lists of c3 / c4 (quick look at kernel durations)
cd $GRAFT_REPO_ROOT
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for c in ${1:-c3 c4}; do
  A="--only $c --only-step --no-graph --steps 3 --warmup 3 --repeats 1 --config-repeats 1"
  python bench.py $A > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics $M --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_tmp_$c.csv python bench.py $A > gpurun_out/ncu_$c.log 2>&1
  python profiles/summarize_launches.py $c=gpurun_out/r02_launches_tmp_$c.csv:6:1 -o /tmp/k.json | grep -A8 "^$c"
done
