set -x
CMD="python bench.py --model EPiC --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline"
$CMD > gpurun_out/plain_epic.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:epic_tile_kernel -s 2 -c 1 -o gpurun_out/prof_epic $CMD > gpurun_out/ncu_epic.log 2>&1
ls -la gpurun_out/prof_epic.ncu-rep
