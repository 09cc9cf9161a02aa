#!/usr/bin/env python
"""Static SASS evidence for profiles/: per hot kernel of libcldet.so, how often the Blackwell-specific instructions occur
(UBLKCP = cp.async.bulk of the TMA unit, SYNCS = mbarrier, UCGABAR = cluster barrier, FFMA2/FMUL2/FADD2 = packed fp32x2,
LDG/STG.256 = 256-bit global accesses, REDG...SYS = system-scope reduction, REDUX = warp-wide reduction).  Runs without a GPU
(cuobjdump).

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'cl_object_detection_b200', 'libcldet.so')
WANT = ['focal_loss_kernel<8, true, false, true, false>', 'focal_loss_kernel<8, true, false, true, true>', 'focal_loss_head_kernel',
        'decode_filter_kernel<4>', 'decode_filter_head_kernel', 'select_fused_kernel', 'radix_sort_kernel', 'nms_mask_kernel',
        'nms_resolve_stream_kernel', 'nms_fused_kernel', 'rank_sort_kernel', 'gt_scatter_kernel', 'focal_reweight_kernel<8, true, false, false>']
PAT = re.compile(r'\b(UBLKCP|UTMALDG|UTMASTG|SYNCS\.\S+|FFMA2|FMUL2|FADD2|LDG\.E\.\S*256|STG\.E\.\S*256|LDG\.E\.\S*128|STG\.E\.\S*128|'
                 r'ACQBULK|UCGABAR\S*|REDG\S*|REDUX\S*|MATCH\S*|ATOMS\S*|BAR\.SYNC\S*|MUFU\.EX2|MUFU\.RCP|MUFU\.LG2|FFMA|FMUL|FADD)\b')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    fn, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            continue
        if fn:
            for t in PAT.findall(line):
                t = re.sub(r'LDG\.E\.\S*256', 'LDG.256', t)
                t = re.sub(r'STG\.E\.\S*256', 'STG.256', t)
                t = re.sub(r'LDG\.E\.\S*128', 'LDG.128', t)
                t = re.sub(r'STG\.E\.\S*128', 'STG.128', t)
                t = re.sub(r'SYNCS\.\S+', 'SYNCS(mbarrier)', t)
                counts[fn][t] += 1
    names = subprocess.run(['c++filt'], input='\n'.join(counts), capture_output=True, text=True).stdout.split('\n')
    dem = dict(zip(counts, names))
    print('# cuobjdump -sass cl_object_detection_b200/libcldet.so (sm_100a): static instruction counts of the hot kernels\' SASS')
    print('# UBLKCP = cp.async.bulk (the TMA unit\'s bulk copy, global<->shared), SYNCS = mbarrier operations, UCGABAR* = thread-block-cluster')
    print('# barrier, FFMA2/FMUL2/FADD2 = packed fp32x2 arithmetic (sm_100), LDG/STG.256 = 256-bit global accesses (sm_100),')
    print('# REDG...SYS = system-scope reduction (peer arrival counter), REDUX = warp-wide OR (the NMS chain\'s parallel rounds), MUFU.* = special-function unit.')
    print('# Regenerate: tools/sass_evidence.py')
    for fn, c in sorted(counts.items(), key=lambda kv: dem[kv[0]]):
        d = dem[fn]
        if any(w in d for w in WANT):
            print('%s\n    %s' % (d[:150], '  '.join('%s=%d' % kv for kv in sorted(c.items()))))


if __name__ == '__main__':
    sys.exit(main())
