# round-1d evidence run: parity tests, smoke, headline bench, other encoders, ncu captures of the step kernel and the tile kernel
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -2 gpurun_out/bench_r01d.err
timeout 300 python bench.py --model FusedParticleFormer --no-cpu-baseline --no-step-roofline > gpurun_out/bench_r01d_fused.json 2>> gpurun_out/bench_r01d.err
timeout 300 python bench.py --model EPiC --no-cpu-baseline --no-step-roofline > gpurun_out/bench_r01d_epic.json 2>> gpurun_out/bench_r01d.err
CMDS="python tools/step_rate.py"
timeout 300 $CMDS > gpurun_out/plain_r01d_step.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:hybrid_step_prod -s 3 -c 1 -f -o gpurun_out/prof_r01d_step $CMDS > gpurun_out/ncu_r01d_step.log 2>&1
CMD="python bench.py --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline"
timeout 300 $CMD > gpurun_out/plain_r01d.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01d.csv $CMD > gpurun_out/ncu_r01d2.log 2>&1
timeout 300 $CMD > gpurun_out/plain_r01d2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01d $CMD > gpurun_out/ncu_r01d.log 2>&1
ls -la gpurun_out | tail -12
