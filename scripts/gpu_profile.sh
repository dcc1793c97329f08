#!/bin/bash
# plain run, then (same command line) the ncu launch list, then one `--set full` capture of selected layers of a step.
# Keep captures small: gpurun only merges gpurun_out/ back when it is < 64 MiB.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-roofline --no-u8 --no-other-precision"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit=$?"
LAYERS="${@:-model.0 model.1.dw model.2.dw model.8.dw cpm.trunk.0.dw model.1.pw model.3.pw model.8.pw cpm.align initial_stage.trunk.0 refinement_stages.0.trunk.4.trunk.1 refinement_stages.0.trunk.0.trunk.1+1.initial initial_stage.heads.fused refinement_stages.0.heads.fused cpm.trunk.0.pw postproc}"
python scripts/prof_layers.py $LAYERS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_layers python scripts/prof_layers.py $LAYERS > gpurun_out/ncu_full.log 2>&1
echo "full capture exit=$?"
du -sh gpurun_out
