( time timeout 900 python -m pytest tests/test_backward_teacher_forced_gpu.py -x -q -m gpu -s ) 2>&1 | tail -40
