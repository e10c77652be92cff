# quick iteration job: parity tests of the tile kernels, short benches, stage traces
set -x
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for M in ParticleFormer FusedParticleFormer; do
  timeout 300 python bench.py --model $M --steps 5 --warmup 3 --no-cpu-baseline --no-step-roofline 2>gpurun_out/iter_$M.err | tee gpurun_out/iter_$M.json | cut -c1-330
  MMF_TRACE=gpurun_out/iter_trace_$M.txt timeout 120 python tools/tf_trace.py $M > /dev/null 2>&1
done
timeout 300 python bench.py --model ParticleFormer --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-step-roofline 2>/dev/null | cut -c1-330
timeout 300 python bench.py --model EPiC --steps 5 --warmup 3 --no-cpu-baseline --no-step-roofline 2>/dev/null | cut -c1-330
timeout 300 python bench.py --model EPiC --batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-step-roofline 2>/dev/null | cut -c1-330
MMF_TRACE=gpurun_out/iter_trace_EPiC.txt timeout 120 python tools/epic_trace.py > /dev/null 2>&1
