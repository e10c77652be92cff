timeout 300 python -m pytest tests/test_gpu_epic.py -q -x 2>&1 | tail -2
timeout 300 python bench.py --model EPiC --steps 5 --warmup 3 --no-cpu-baseline --no-step-roofline 2>/dev/null | cut -c1-230
timeout 300 python bench.py --model EPiC --batch 4096 --steps 3 --warmup 3 --no-cpu-baseline --no-step-roofline 2>/dev/null | cut -c1-230
MMF_TRACE=gpurun_out/iter_trace_EPiC.txt timeout 120 python tools/epic_trace.py > /dev/null 2>&1
