"""Build libcldet.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

Called by __graft_entry__.build() and by `python -m cl_object_detection_b200.build`.  nvcc cross-compiles
without a GPU.  The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
INCLUDE = os.path.join(ROOT, 'include')
LIB = os.path.join(PKG, 'libcldet.so')
OBJ_DIR = os.path.join(ROOT, 'build', 'cldet')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-I' + INCLUDE, '-I' + CSRC]
# Units whose fp32 results decide integer outputs (assignment, NMS keep, decoded boxes) are built WITHOUT
# FMA contraction so that every op rounds exactly like the reference's un-fused ATen ops.
SOURCES = {
    'cldet_assign.cu': ['-fmad=false'],
    'cldet_detect.cu': ['-fmad=false'],
    'cldet_loss.cu': [],
    'cldet_loss_logits.cu': [],
    'cldet_distill.cu': [],
}


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libcldet.so cannot be built')


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')] + \
        [os.path.join(INCLUDE, 'cldet.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + ARCH + COMMON + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s' % (' '.join(cmd), out))
    tmp = LIB + '.tmp.%d' % os.getpid()
    cmd = [nvcc] + ARCH + ['-shared', '-o', tmp] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed: %s\n%s' % (' '.join(cmd), r.stdout))
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
