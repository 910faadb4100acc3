"""Compile the CUDA sources in `mpqe_b200/csrc` into the in-tree C-ABI library `mpqe_b200/_C/libmpqe_b200.so`.

    python -m mpqe_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.environ.get('MPQE_BUILD_DIR') or os.path.join(HERE, '_C')   # (debug builds go to a directory of their own)
LIB_PATH = os.path.join(OUT_DIR, 'libmpqe_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--use_fast_math=false',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr']
FLAGS.remove('--use_fast_math=false')  # precise math only: fp32 parity with the reference is a requirement
FLAGS += os.environ.get('MPQE_NVCC_FLAGS', '').split()  # e.g. -DMPQE_TC_STATS for the debug cycle accounting


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _fingerprint():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/mpqe_b200.h']:
        path = os.path.normpath(os.path.join(CSRC, f))
        if os.path.isfile(path):
            h.update(f.encode())
            h.update(open(path, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, 'build.stamp')
    fp = _fingerprint()
    if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp) and open(stamp).read() == fp:
        return LIB_PATH
    if not os.path.isfile(NVCC):
        raise RuntimeError('nvcc not found at %s and no prebuilt %s' % (NVCC, LIB_PATH))
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('--- %s\n%s\n' % (os.path.basename(src), out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    cmd = [NVCC, '-shared', '-o', LIB_PATH] + objs + ['-cudart', 'static']
    subprocess.run(cmd, check=True)
    with open(stamp, 'w') as f:
        f.write(fp)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
