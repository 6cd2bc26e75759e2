"""In-tree build of libvqnerf_b200.so (nvcc, sm_100a only).

`python -m vqnerf_release_b200.build` compiles every .cu under csrc/ with
`-gencode arch=compute_100a,code=sm_100a -lineinfo` and links them into
vqnerf_release_b200/libvqnerf_b200.so.  nvcc cross-compiles without a GPU; the
built .so travels to the GPU box with the repo snapshot (it is git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libvqnerf_b200.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
    '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr',
] + os.environ.get('VQN_EXTRA_NVCC_FLAGS', '').split()      # e.g. -DVQN_TC_TRACE for benchmarks/tc_trace.py


def _nvcc() -> str:
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest(path: str) -> str:
    h = hashlib.sha256()
    for dep in [path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h'))] + \
            [os.path.join(HERE, '..', 'include', 'vqnerf_b200.h')]:
        with open(dep, 'rb') as fh:
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    objs, jobs = [], []
    for src in _sources():
        sp = os.path.join(CSRC, src)
        obj = os.path.join(BUILD, src[:-3] + '.o')
        stamp = obj + '.sha'
        dg = _digest(sp)
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.exists(stamp)
                and open(stamp).read() == dg):
            continue
        jobs.append((sp, obj, stamp, dg))

    def compile_one(job):
        sp, obj, stamp, dg = job
        cmd = [nvcc] + NVCC_FLAGS + ['-c', sp, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = obj + '.log'
        with open(log, 'w') as fh:
            fh.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s' % (sp, r.stderr[-4000:]))
        with open(stamp, 'w') as fh:
            fh.write(dg)
        return log

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            logs = list(ex.map(compile_one, jobs))
        if verbose:
            for lg in logs:
                print(open(lg).read())
    if jobs or not os.path.exists(LIB) or force:
        # no -lcuda: the one driver entry point the library needs (cuTensorMapEncodeTiled) is resolved at run time through
        # cudaGetDriverEntryPoint, so the .so loads on machines without a driver (build checks, symbol tests)
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                     '-Xcompiler', '-fPIC']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stderr[-4000:])
    return LIB


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
