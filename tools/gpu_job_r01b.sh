set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err; tail -2 gpurun_out/bench_r01b.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01b_ref.json 2>> gpurun_out/bench_r01b.err
CMD="python bench.py --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline"
$CMD > gpurun_out/plain_tile.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tile.csv $CMD > gpurun_out/ncu_tile1.log 2>&1
$CMD > gpurun_out/plain_tile2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -o gpurun_out/prof_tile $CMD > gpurun_out/ncu_tile2.log 2>&1
ls -la gpurun_out | tail -12
