"""Builds v2ce_toolbox_b200/libv2ce_b200.so in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the
GPU box with the gpurun snapshot.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libv2ce_b200.so')
STAMP = os.path.join(HERE, '.libv2ce_b200.stamp')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr',
              '-cudart', 'static']


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _digest():
    h = hashlib.sha256()
    for p in _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))) + \
            [os.path.join(os.path.dirname(HERE), 'include', 'v2ce_b200.h')]:
        with open(p, 'rb') as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_library(force=False, verbose=False):
    """Compile every .cu under csrc/ (one nvcc process per file, in parallel) and link."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == digest:
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f'--- {os.path.basename(src)} ---\n{out}\n')
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-cudart', 'static', '-o', LIB] + objs
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
