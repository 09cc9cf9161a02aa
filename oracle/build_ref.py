"""TEST / BASELINE INFRASTRUCTURE -- never imported by the product path.

Recipe for `oracle/_ref/`: the reference's OWN implementation of the path, built from the sources where they lie under
/root/reference (read-only) into byte-code only.  The reference is pure Python, so "compiling" it means py_compile: the
modules the path lives in (retinanet/losses.py FocalLoss + calc_iou, retinanet/anchors.py, retinanet/utils.py
BBoxTransform / ClipBoxes, retinanet/model.py ResNet.predict and the print helper it imports) become byte-code files `oracle/_ref/retinanet/<module>.bytecode` (the .pyc format under another
extension: snapshot tools that skip `*.pyc` as caches would drop them; oracle/ref_runner.py imports them with
importlib's SourcelessFileLoader).  No reference SOURCE is copied
into the repo; oracle/_ref/ is git-ignored (it is a build output) but not gpurun-ignored, so it travels to the GPU box like
libcldet.so does.  /root/reference itself does not exist there and nothing reads it at run time.

Used by: bench.py's reference arm / cpu_baseline leg (kind = "reference"), tests/test_reference_snapshot.py (the numpy
restatement against the live reference on fresh random inputs, wherever the snapshot exists).

    python -m oracle.build_ref          # in the build container; __graft_entry__.build() calls build_ref()
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
REFERENCE = os.environ.get('CLDET_REFERENCE', '/root/reference')
EXT = '.bytecode'
MODULES = ('retinanet/__init__.py', 'retinanet/losses.py', 'retinanet/anchors.py', 'retinanet/utils.py',
           # the eval half: ResNet.predict lives in model.py, which imports preprocessing.debug (a print helper)
           'retinanet/model.py', 'preprocessing/__init__.py', 'preprocessing/debug.py')


def build_ref(force=False):
    """Returns the snapshot directory, or None when there is no reference tree to build from (the GPU box: the prebuilt
    snapshot that travelled with the repo is used as is)."""
    if not os.path.isdir(REFERENCE):
        return OUT if available() else None
    stamp = os.path.join(OUT, 'BUILT_FOR')
    tag = '%s %s' % (sys.implementation.cache_tag, sys.version.split()[0])
    if not force and os.path.exists(stamp) and open(stamp).read().strip() == tag and \
            all(os.path.exists(os.path.join(OUT, m[:-3] + EXT)) for m in MODULES):
        return OUT
    for m in MODULES:
        src = os.path.join(REFERENCE, m)
        dst = os.path.join(OUT, m[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        py_compile.compile(src, cfile=dst, dfile='<reference>/' + m, doraise=True, optimize=0)
    with open(stamp, 'w') as f:
        f.write(tag + '\n')
    return OUT


def available():
    return os.path.exists(os.path.join(OUT, 'retinanet', 'losses' + EXT))


if __name__ == '__main__':
    print(build_ref(force='--force' in sys.argv))
