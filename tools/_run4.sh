timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_ops_b256_gpu.py tests/test_student_gpu.py tests/test_backward_teacher_forced_gpu.py tests/test_parity_configs_gpu.py -x -q -m gpu 2>&1 | tail -2
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-configs --no-e2e --no-cpu-baseline"
$B > gpurun_out/r02w.json 2> gpurun_out/r02w.err
timeout 20 python tools/benchsum.py < gpurun_out/r02w.json
