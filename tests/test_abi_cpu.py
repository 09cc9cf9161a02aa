"""CPU checks of the C-ABI boundary: the library builds/loads and exports every symbol include/cldet.h declares,
the ctypes table covers the header, and the product path refuses to run without CUDA (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

import cl_object_detection_b200 as cld
from cl_object_detection_b200 import _lib
from oracle import head_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'cldet.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cldet_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from cl_object_detection_b200.build import build_library
    build_library()
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), 'libcldet.so does not export %s' % n
    assert set(names) == set(_lib.SIGNATURES), 'ctypes table and header disagree'
    assert not _lib._PENDING
    assert lib.cldet_abi_version() == 1
    assert lib.cldet_status_string(0) == b'ok'


def test_num_anchors_matches_oracle_without_gpu():
    for hw in [(512, 512), (800, 1333), (1333, 1333), (33, 70), (608, 1024), (1, 1)]:
        assert cld.num_anchors(*hw) == O.num_anchors(*hw)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    assert lib.cldet_anchors(0, 10, None, None) == 1
    assert lib.cldet_focal_loss_workspace_bytes(0, 10) == 0
    assert lib.cldet_focal_loss_workspace_bytes(2, 1000) > 0
    assert lib.cldet_iou_assign(None, 10, None, 1, 1, 1, None, None, None, None, None, None) == 1


def test_no_cpu_fallback():
    fl = cld.FocalLoss()
    p = cld.HeadParams()
    with pytest.raises(RuntimeError, match='CUDA'):
        fl(torch.rand(1, 9, 2), torch.rand(1, 9, 4), torch.rand(1, 9, 4), -torch.ones(1, 1, 5), 0, p)
    with pytest.raises(RuntimeError, match='CUDA'):
        cld.calc_iou(torch.rand(3, 4), torch.rand(2, 4))


def _imported_modules(path):
    """Every module name a Python source imports (absolute and relative), from its AST."""
    import ast
    tree = ast.parse(open(path).read(), filename=path)
    names = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Import):
            names += [a.name for a in node.names]
        elif isinstance(node, ast.ImportFrom):
            names.append(('.' * node.level) + (node.module or ''))
            names += [('.' * node.level) + (node.module + '.' if node.module else '') + a.name for a in node.names]
        elif isinstance(node, ast.Call) and getattr(node.func, 'id', getattr(node.func, 'attr', '')) in ('__import__', 'import_module'):
            names += [a.value for a in node.args[:1] if isinstance(a, ast.Constant) and isinstance(a.value, str)]
    return names


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it (AST scan of every import statement,
    __import__ / import_module call), and no native source may include anything from oracle/."""
    pkg = os.path.join(ROOT, 'cl_object_detection_b200')
    seen = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith('.py'):
                seen += 1
                for name in _imported_modules(path):
                    parts = name.lstrip('.').split('.')
                    assert 'oracle' not in parts, '%s imports %s' % (path, name)
            elif f.endswith(('.cu', '.cuh', '.h', '.cpp')):
                for line in open(path):
                    if line.lstrip().startswith('#include'):
                        assert 'oracle' not in line, '%s: %s' % (path, line)
    assert seen >= 8


def test_params_translation_matches_reference_error_behaviour():
    from cl_object_detection_b200.params import to_loss_params
    p = cld.HeadParams([0, 15])
    lp = to_loss_params(p, 1, 16)
    assert lp.incremental == 1 and lp.past_class_num == 15 and abs(lp.alpha - 0.25) < 1e-7
    p['decrease_positive'] = None
    with pytest.raises(TypeError):
        to_loss_params(p, 1, 16)
    p['decrease_positive_by_IOU'] = True
    assert to_loss_params(p, 1, 16).decrease_positive_by_iou == 1


def test_a9_pseudo_label_format_mirrors_reference_collater():
    """Host-side format helpers (CPU): merged pseudo rows + -1 padding exactly as the reference's collater emits them."""
    from tests.helpers import load
    g = load('a9_pseudo_labels')
    annots = [g['annot0'], g['annot1'], g['annot2']]
    out = cld.collate_annotations(annots)
    assert out.dtype == torch.float32 and np.array_equal(out.numpy(), g['collated'])
    assert np.array_equal(cld.collate_annotations([np.zeros((0, 5))]).numpy(), g['collated_empty'])
    real = np.array([[10, 20, 30, 40, 15]], dtype=np.float64)
    pseudo = np.array([[5, 6, 7, 8, 1], [1, 2, 3, 4, 0]], dtype=np.float64)
    assert np.array_equal(cld.merge_pseudo_labels(real, pseudo), O.merge_pseudo_labels(real, pseudo))


def test_torch_op_layer_builds_loads_and_registers():
    """The C++ custom-op layer (csrc/cldet_torch.cpp) builds against this torch, links to libcldet.so and registers its
    schemas; without a GPU its ops refuse CPU tensors (no fallback)."""
    from cl_object_detection_b200.build import build_ops
    build_ops()
    ops = cld.load_ops()
    assert ops.abi_version() == 1
    for name in ('focal_loss', 'detect', 'batched_nms'):
        assert hasattr(ops, name)
    schema = str(torch.ops.cldet.focal_loss.default._schema)
    assert 'Tensor cls' in schema and 'int[] peer' in schema and schema.endswith('-> Tensor[]')
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.batched_nms(torch.rand(3, 4), torch.rand(3), None, 0.5, 0, 100000)
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.detect(torch.rand(1, 9, 2), torch.rand(1, 9, 4), torch.rand(1, 9, 4), 8, 8, True, 0.05, 0.5, 0, 0, 100000, 0)
    with pytest.raises(RuntimeError, match='CUDA'):
        cld.detect.detect_batch(torch.rand(1, 9, 2), torch.rand(1, 9, 4), torch.rand(1, 9, 4), 8, 8)
