#!/usr/bin/env python
"""Build the in-tree CUDA library (sm_100a only) and, optionally, list its exports.

    python build.py            # incremental: per-file objects under build/, then link
    python build.py --force    # rebuild everything
    python build.py --verbose  # add -Xptxas -v

Output: ubisoft-laforge-msmd_b200/libmsmd_b200.so (git-ignored; travels to the GPU box).
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'ubisoft-laforge-msmd_b200')
CSRC = os.path.join(PKG, 'csrc')
OUT = os.path.join(PKG, 'libmsmd_b200.so')
OBJ = os.path.join(ROOT, 'build')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-I', os.path.join(ROOT, 'include')]
FLAGS += os.environ.get('MSMD_EXTRA_NVCC_FLAGS', '').split()     # e.g. -DMSMD_ROW0_TRACE for the timeline probes


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))


def headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hs.append(os.path.join(ROOT, 'include', 'msmd_b200.h'))
    return max(os.path.getmtime(h) for h in hs)


def compile_one(src, verbose):
    obj = os.path.join(OBJ, src[:-3] + '.o')
    cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', os.path.join(CSRC, src), '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False, quiet=False):
    os.makedirs(OBJ, exist_ok=True)
    hm = headers_mtime()
    todo = []
    for s in sources():
        obj = os.path.join(OBJ, s[:-3] + '.o')
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(hm, os.path.getmtime(os.path.join(CSRC, s))):
            todo.append(s)
    ok = True
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, rc, log in ex.map(lambda s: compile_one(s, verbose), todo):
            if rc != 0 or (verbose and log.strip()):
                print(f'--- {src} (rc={rc})\n{log}')
            ok &= rc == 0
    if not ok:
        raise RuntimeError('nvcc failed')
    objs = [os.path.join(OBJ, s[:-3] + '.o') for s in sources()]
    if todo or not os.path.exists(OUT):
        cmd = [NVCC, '-shared', '-o', OUT] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    if not quiet:
        print(f'built {OUT} ({len(todo)} file(s) recompiled)')
    return OUT


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    build(a.force, a.verbose)
