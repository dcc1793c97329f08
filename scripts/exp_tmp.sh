LAYERS="initial_stage.trunk.0 refinement_stages.0.trunk.0.trunk.1 model.8.pw model.8.dw postproc"
python scripts/prof_layers.py $LAYERS > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_layers_f python scripts/prof_layers.py $LAYERS > gpurun_out/ncu_full_f.log 2>&1
echo "full capture exit=$?"
