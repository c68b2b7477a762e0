timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_student_gpu.py tests/test_augment_gpu.py -x -q -m gpu 2>&1 | tail -2
VPD_K1_ARITH=0 timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k assemble 2>&1 | tail -2
VPD_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum --clock-control none -k regex:'assemble_pad8' -c 3 python tools/prof_step.py 2 2>&1 | grep -E "assemble_pad8|gpu__time|wavefronts|inst_executed"
