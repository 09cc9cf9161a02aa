#!/usr/bin/env python
"""Build an A/B variant of libcldet.so with extra -D flags (experiments only; the shipped library is built by
cl_object_detection_b200/build.py).  python tools/build_variant.py NAME -DCLDET_LOSS_PRELOAD ... -> build/variants/libcldet_NAME.so
Select it at run time with CLDET_LIBRARY=build/variants/libcldet_NAME.so."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cl_object_detection_b200 import build as B  # noqa: E402


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out_dir = os.path.join(ROOT, 'build', 'variants')
    obj_dir = os.path.join(out_dir, name)
    os.makedirs(obj_dir, exist_ok=True)
    procs, objs = [], []
    for src, extra in B.SOURCES.items():
        obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
        cmd = [B._nvcc()] + B.ARCH + B.COMMON + extra + flags + ['-Xptxas', '-v', '-c', os.path.join(B.CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode:
            raise SystemExit(out)
        if src == 'cldet_loss.cu':
            lines = out.splitlines()
            for i, ln in enumerate(lines):
                if 'focal_loss_kernelILi8ELb1ELb0ELb1ELb0' in ln and 'Compiling' in ln:
                    print('\n'.join(lines[i:i + 4]))
    lib = os.path.join(out_dir, 'libcldet_%s.so' % name)
    subprocess.run([B._nvcc()] + B.ARCH + ['-shared', '-o', lib] + objs, check=True)
    print(lib)


if __name__ == '__main__':
    main()
