"""Build libvpd_b200.so in-tree with nvcc for sm_100a (no torch headers, no JIT).

    python -m vpd_b200.build [--force]

The shared library exposes only the C ABI of include/vpd_b200.h; cudart is
linked statically so the .so loads on a box without a CUDA driver (symbol
checks) and on the GPU box (compute).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libvpd_b200.so')
OBJ = os.path.join(HERE, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
         '-Xptxas', '-v']


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _deps_mtime():
    m = 0
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    hdr_m = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)
                if f.endswith(('.h', '.cuh')))
    hdr_m = max(hdr_m, os.path.getmtime(os.path.join(HERE, '..', 'include', 'vpd_b200.h')))
    procs = []
    objs = []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + '.o')
        objs.append(o)
        if (not force and os.path.exists(o)
                and os.path.getmtime(o) >= max(os.path.getmtime(s), hdr_m)):
            continue
        cmd = [NVCC] + FLAGS + ['-c', s, '-o', o]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append('== {} ==\n{}'.format(src, out))
        if p.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed for {}:\n{}\n'.format(src, out))
    with open(os.path.join(OBJ, 'ptxas.log'), 'a' if not force else 'w') as fp:
        fp.write('\n'.join(log))
    if failed:
        raise RuntimeError('vpd_b200: nvcc build failed')
    if verbose:
        print('\n'.join(log))
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-cudart', 'static', '-Xcompiler', '-fPIC']
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
