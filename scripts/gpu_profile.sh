#!/bin/bash
# plain run, then (same command line) the ncu launch list and full captures of the top kernels.
# Keep captures small: gpurun only merges gpurun_out/ back when it is < 64 MiB.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit=$?"
i=0
for spec in "$@"; do   # spec = regex:skip:count:name
  IFS=: read -r rx skip cnt name <<< "$spec"
  $CMD > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"$rx" -s "$skip" -c "$cnt" -f -o gpurun_out/prof_$name $CMD > gpurun_out/ncu_full_$name.log 2>&1
  echo "full capture $name exit=$?"
done
du -sh gpurun_out
