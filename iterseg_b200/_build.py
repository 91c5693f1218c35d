"""In-tree build of libiterseg_b200.so (nvcc, sm_100a only).

`python -m iterseg_b200._build` or `__graft_entry__.build()`.  nvcc
cross-compiles without a GPU; the resulting .so is git-ignored but travels with
the working tree to the GPU box.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(CSRC, '_obj')
LIB = os.path.join(PKG, 'libiterseg_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
              '-std=c++17', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'),
              '-I', CSRC] + os.environ.get('ISG_NVCC_EXTRA', '').split()


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = os.environ.get('NVCC', 'nvcc')
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(glob.glob(os.path.join(CSRC, '*.cu')))
    headers = sorted(glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(CSRC, '*.h'))
                     + glob.glob(os.path.join(ROOT, 'include', '*.h')))
    jobs = []
    objs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError('nvcc failed for ' + cmd[-3])
    if jobs or force or _newer(LIB, objs):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
