# N fresh processes of tools/tile_stress.py (the first launches of a process are where a cold-start race shows):
#   bash tools/gpu_job_fresh_env.sh <N> [VAR=value ...]      e.g.  ... 8 STRESS_JETS=pair
# A library built with MMF_EXTRA_NVCC="-DMMF_PROD_DIAG=1" python multimodal-flows_b200/build.py --force prints, when a
# process exits, which warp timed out on which barrier.
N=$1; shift
fails=0
for i in $(seq 1 $N); do
  env "$@" timeout 120 python tools/tile_stress.py 2 > /tmp/fr.log 2>&1 || { fails=$((fails+1)); echo "--- process $i"; grep "repeat [0-9]* nsteps\|timed-out\|  cta " /tmp/fr.log | cut -c1-140; }
done
echo "== env $*: $fails of $N fresh processes failed"
