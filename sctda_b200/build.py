"""Builds sctda_b200/csrc/libsctda_b200.so (sm_100a only) with nvcc, in tree.

No torch, no JIT cache: the .so travels with the repo snapshot to the GPU box.
`python -m sctda_b200.build [--force]`.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, 'libsctda_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include')]


# sources whose float64 arithmetic must not be contracted into FMAs by the compiler (explicit
# __fma_rn only): their results are compared bit for bit with the scalar oracle
NO_FMAD = ('cluster.cu', 'dist.cu')


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link the shared library.  Returns its path."""
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(ROOT, 'include', 'sctda_b200.h'))
    objs = []
    procs = []
    for src in sources():
        obj = src[:-3] + '.o'
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            extra = ['-fmad=false'] if os.path.basename(src) in NO_FMAD else []   # fixed operation order
            cmd = [nvcc] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), out))
        if verbose:
            sys.stderr.write(out)
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-ldl']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError('link failed: %s\n%s' % (' '.join(cmd), r.stdout.decode()))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
