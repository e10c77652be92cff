set -x
timeout 300 python -m pytest tests/test_gpu_epic.py -q 2>&1 | tail -2
python bench.py --model EPiC --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | cut -c1-200
CMD="python bench.py --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline"
$CMD > gpurun_out/plain_tile3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -o gpurun_out/prof_tile2 $CMD > gpurun_out/ncu_tile3.log 2>&1
ls -la gpurun_out/prof_tile2.ncu-rep
