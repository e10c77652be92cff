# ncu evidence for the tile kernel (run AFTER the same command exited 0 without ncu)
set -x
CMD="python bench.py --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline"
$CMD > gpurun_out/plain_r01c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -o gpurun_out/prof_r01c $CMD > gpurun_out/ncu_r01c.log 2>&1
$CMD > gpurun_out/plain_r01c2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01c.csv $CMD > gpurun_out/ncu_r01c2.log 2>&1
ls -la gpurun_out/prof_r01c.ncu-rep
