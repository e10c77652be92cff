# final-build evidence: GPU suite twice (flakiness), full ncu capture of the plain tile kernel of the bench command
set -x
for i in 1 2; do timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -1; done
CMD="python bench.py --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline --no-extras"
$CMD > gpurun_out/r3g_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -o /tmp/r3g_tile $CMD > gpurun_out/r3g_ncu.log 2>&1
ncu -i /tmp/r3g_tile.ncu-rep --page raw --csv > gpurun_out/r3g_tile_raw.csv
ncu -i /tmp/r3g_tile.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r3g_tile_source.csv 2>/dev/null
ls -la gpurun_out/r3g_*
