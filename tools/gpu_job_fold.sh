# LayerNorm fold + softmax without the row maximum: parity suite, then rates with / without the maximum
set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-step-roofline --no-extras"
MMF_TILE_DEBUG=1 $B 2>&1 | grep -E "score bound|\"value\"" | cut -c1-180
MMF_TILE_SOFTMAX_MAX=1 MMF_TILE_DEBUG=1 $B 2>&1 | grep -E "score bound|\"value\"" | cut -c1-180
$B --model FusedParticleFormer 2>/dev/null | cut -c1-180
$B --batch 4096 --steps 2 2>/dev/null | cut -c1-180
$B --dense --steps 2 2>/dev/null | cut -c1-180
MMF_TILE_SOFTMAX_MAX=1 timeout 600 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_sampler.py tests/test_gpu_parity_r2.py -q -x 2>&1 | tail -2
