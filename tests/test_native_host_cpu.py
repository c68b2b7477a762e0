"""Host-side native logic that needs no GPU: built with nvcc (cross-compiles here) and run on the
CPU. Currently the division-free tile decode of the conv kernels (FastDiv)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')


@pytest.mark.skipif(not os.path.exists(NVCC) and shutil.which('nvcc') is None, reason='nvcc not available')
def test_fastdiv_matches_integer_division(tmp_path):
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which('nvcc')
    exe = str(tmp_path / 'fastdiv_check')
    subprocess.check_call([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-O2', '-std=c++17',
                           '--expt-relaxed-constexpr', '-I', os.path.join(ROOT, 'vpd_b200', 'csrc'),
                           os.path.join(ROOT, 'tests', 'native', 'fastdiv_check.cu'), '-o', exe,
                           '-cudart', 'static'], stderr=subprocess.DEVNULL)
    out = subprocess.check_output([exe]).decode()
    assert out.startswith('ok '), out
