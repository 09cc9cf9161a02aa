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
OPS_LIB = os.path.join(PKG, '_cldet_torch.so')          # the torch custom-op layer (csrc/cldet_torch.cpp) over the C ABI
OPS_SRC = os.path.join(CSRC, 'cldet_torch.cpp')
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


def ops_is_stale():
    if not os.path.exists(OPS_LIB):
        return True
    t = os.path.getmtime(OPS_LIB)
    return any(os.path.getmtime(d) > t for d in (OPS_SRC, os.path.join(INCLUDE, 'cldet.h')))


def build_ops(force=False, verbose=False):
    """Compile the torch custom-op layer (host C++ only: it holds no kernel) against this interpreter's torch and link it to
    libcldet.so next to it ($ORIGIN rpath).  Loaded with torch.ops.load_library, so it needs no Python headers."""
    build_library(force=False)
    if not force and not ops_is_stale():
        return OPS_LIB
    import torch
    from torch.utils import cpp_extension as ce
    cxx = os.environ.get('CXX') or shutil.which('g++') or 'g++'
    tlib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    cuda_home = ce.CUDA_HOME or '/usr/local/cuda'
    inc = ['-I' + p for p in ce.include_paths()] + ['-I' + os.path.join(cuda_home, 'include'), '-I' + INCLUDE]
    tmp = OPS_LIB + '.tmp.%d' % os.getpid()
    cmd = [cxx, '-O2', '-std=c++17', '-fPIC', '-shared', '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI),
           '-DTORCH_API_INCLUDE_EXTENSION_H', '-Wno-deprecated-declarations'] + inc + [OPS_SRC, '-o', tmp,
           '-L' + tlib, '-ltorch', '-ltorch_cpu', '-ltorch_cuda', '-lc10', '-lc10_cuda', '-L' + PKG, '-l:libcldet.so',
           '-Wl,-rpath,$ORIGIN', '-Wl,--no-as-needed']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose and r.stdout:
        print(r.stdout)
    if r.returncode != 0:
        raise RuntimeError('building the torch op layer failed: %s\n%s' % (' '.join(cmd), r.stdout))
    os.replace(tmp, OPS_LIB)
    return OPS_LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose='-v' in sys.argv))
    print(build_ops(force='--force' in sys.argv, verbose='-v' in sys.argv))
