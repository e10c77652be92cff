# two-issuer experiment: dead-lock check under the trace build (bounded waits), then parity tests and short benches
set -x
for M in FusedParticleFormer ParticleFormer; do
  MMF_TRACE=gpurun_out/dbg_trace_$M.txt timeout 120 python tools/dbg_one.py $M 256 4 2>&1 | tail -30
done
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for M in ParticleFormer FusedParticleFormer; do
  timeout 300 python bench.py --model $M --steps 5 --warmup 3 --no-cpu-baseline --no-step-roofline 2>gpurun_out/iter_$M.err | tee gpurun_out/iter_$M.json | cut -c1-330
done
