# end-of-round check of HEAD: all GPU parity tests, smoke, fresh-process stress, the default bench line (with rooflines and CPU baseline), the reference arm
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash tools/gpu_job_fresh_env.sh 6 2>&1 | tail -4
( time timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | tail -4; tail -2 gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2>> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final_ref.json
