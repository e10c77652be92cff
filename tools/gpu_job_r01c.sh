# round-1c evidence run: parity tests, smoke, headline bench + reference arm, config sweep, stage traces, ncu captures
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; tail -2 gpurun_out/bench_r01c.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01c_ref.json 2>> gpurun_out/bench_r01c.err
timeout 300 python bench.py --model FusedParticleFormer --no-cpu-baseline --no-step-roofline > gpurun_out/bench_r01c_fused.json 2>> gpurun_out/bench_r01c.err
timeout 300 python bench.py --model EPiC --no-cpu-baseline --no-step-roofline > gpurun_out/bench_r01c_epic.json 2>> gpurun_out/bench_r01c.err
timeout 300 python bench.py --dense --no-cpu-baseline --no-step-roofline > gpurun_out/bench_r01c_dense.json 2>> gpurun_out/bench_r01c.err
timeout 900 python tools/sweep_configs.py all > gpurun_out/sweep_r01c.jsonl 2>> gpurun_out/bench_r01c.err
for M in ParticleFormer FusedParticleFormer; do MMF_TRACE=gpurun_out/trace_r01c_$M.txt timeout 120 python tools/tf_trace.py $M > /dev/null 2>&1; done
MMF_TRACE=gpurun_out/trace_r01c_EPiC.txt timeout 120 python tools/epic_trace.py > /dev/null 2>&1
CMD="python bench.py --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline"
timeout 300 $CMD > gpurun_out/plain_r01c.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01c $CMD > gpurun_out/ncu_r01c.log 2>&1
timeout 300 $CMD > gpurun_out/plain_r01c2.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_r01c2.log 2>&1
CMDE="python bench.py --model EPiC --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline"
timeout 300 $CMDE > gpurun_out/plain_r01c3.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:epic_tile_kernel -s 2 -c 1 -f -o gpurun_out/prof_r01c_epic $CMDE > gpurun_out/ncu_r01c3.log 2>&1
ls -la gpurun_out | tail -20
