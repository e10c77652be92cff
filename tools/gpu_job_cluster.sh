for C in 1 2 4; do
  echo "cluster $C"
  MMF_TILE_CLUSTER=$C timeout 300 python bench.py --model ParticleFormer --steps 5 --warmup 3 --no-cpu-baseline --no-step-roofline 2>/dev/null | cut -c1-200
  MMF_TILE_CLUSTER=$C timeout 300 python bench.py --model FusedParticleFormer --steps 5 --warmup 3 --no-cpu-baseline --no-step-roofline 2>/dev/null | cut -c1-200
done
