timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-configs --no-e2e --no-cpu-baseline"
$B > gpurun_out/r02w.json 2> gpurun_out/r02w.err
timeout 20 python tools/benchsum.py < gpurun_out/r02w.json
