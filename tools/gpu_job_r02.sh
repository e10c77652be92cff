# round-2 evidence run: full GPU suite, smoke, headline bench (+extras), reference arm, launch list and ncu --set full captures
# of the plain and the pair tile kernel, stage traces, cliff
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -c 300 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_reference_arm.json 2>/dev/null
timeout 300 python tools/cliff_rate.py > gpurun_out/r02_cliff.jsonl 2>&1
for M in FusedParticleFormer ParticleFormer; do
  MMF_TRACE=gpurun_out/r02_trace_pair_$M.txt timeout 120 python tools/tf_trace.py $M dense > /dev/null 2>&1
  MMF_TRACE=gpurun_out/r02_trace_plain_$M.txt timeout 120 python tools/tf_trace.py $M > /dev/null 2>&1
done
CMD="python bench.py --steps 2 --warmup 1 --timesteps 20 --no-cpu-baseline --no-step-roofline --no-extras"
timeout 300 $CMD > gpurun_out/r02_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
timeout 300 $CMD > gpurun_out/r02_plain2.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_tile $CMD > gpurun_out/r02_ncu_tile.log 2>&1
CMDD="python bench.py --steps 2 --warmup 1 --timesteps 20 --batch 74 --dense --no-cpu-baseline --no-step-roofline --no-extras"
timeout 300 $CMDD > gpurun_out/r02_plain3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:tf_tile_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_pair $CMDD > gpurun_out/r02_ncu_pair.log 2>&1
ls -la gpurun_out | tail -15
