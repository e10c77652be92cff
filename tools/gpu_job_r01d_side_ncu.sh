# ncu --set full captures of the two side kernels (each after the same command exited 0 without ncu)
set -x
timeout 300 python tools/obs_rate.py > gpurun_out/plain_r01d_obs.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jet_observables_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01d_obs python tools/obs_rate.py > gpurun_out/ncu_r01d_obs.log 2>&1
timeout 300 python tools/source_rate.py > gpurun_out/plain_r01d_src.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:source_fill_kernel -s 3 -c 1 -f -o gpurun_out/prof_r01d_src python tools/source_rate.py > gpurun_out/ncu_r01d_src.log 2>&1
tail -2 gpurun_out/plain_r01d_obs.log | cut -c1-300; tail -1 gpurun_out/plain_r01d_src.log | cut -c1-300
ls -la gpurun_out/prof_r01d_obs.ncu-rep gpurun_out/prof_r01d_src.ncu-rep
