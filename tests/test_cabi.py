"""CPU: the C-ABI shared library loads and exports exactly what include/peagnn.h declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'peagnn.h')


def _declared():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    src = re.sub(r'typedef struct \{.*?\} peagnn_csr_t;', '', src, flags=re.S)
    out = {}
    for m in re.finditer(r'\b(peagnn_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ('', 'void') else len([a for a in args.split(',') if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope='module')
def lib():
    from graph_recsys_benchmark_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    from graph_recsys_benchmark_b200 import _lib
    decl = _declared()
    assert len(decl) >= 25
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, n in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == n, name


def test_every_declared_symbol_is_exported(lib):
    for name in _declared():
        assert getattr(lib, name) is not None


def test_plain_c_calls(lib):
    assert lib.peagnn_version() >= 100
    assert lib.peagnn_last_error() is not None
    assert lib.peagnn_partial_floats(10, 64, 1) >= 10 * 64
    assert lib.peagnn_wgrad_workspace_floats(1000, 64, 64) >= 64 * 64 + 64
    assert lib.peagnn_bpr_workspace_floats(1024, 16) > 2 * 1024 * 68
    assert lib.peagnn_fuse_workspace_floats(1000, 9, 16) >= 9 * 16


def test_csr_struct_layout_matches_header():
    from graph_recsys_benchmark_b200 import _lib
    # 2 ptr, 4 int32, 2 ptr, 1 int32 (+pad), 3 ptr, 1 ptr, int64, 2 int32, 2 ptr (the optional filters) on LP64
    assert ctypes.sizeof(_lib.CsrView) == 8 * 2 + 4 * 4 + 8 * 2 + 8 + 8 * 3 + 8 + 8 + 8 + 8 * 2


def test_group_problem_structs_match_header():
    """The problem tables of the grouped launches (peagnn_linear_problem_t / peagnn_wgrad_problem_t): field order and
    sizes of the ctypes mirrors follow the header's typedefs."""
    from graph_recsys_benchmark_b200 import _lib
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    for typedef, mirror in (('peagnn_linear_problem_t', _lib.LinearProblem), ('peagnn_wgrad_problem_t', _lib.WgradProblem)):
        body = re.search(r'typedef struct \{([^{}]*)\} %s;' % typedef, src).group(1)
        fields = [re.split(r'[\s\*]+', d.strip())[-1] for d in body.split(';') if d.strip()]
        assert fields == [name for name, _ in mirror._fields_], typedef
        assert ctypes.sizeof(mirror) == 8 * len(fields)              # pointers and int64 only: no padding on LP64
    assert int(re.search(r'#define PEAGNN_MAX_GROUP (\d+)', src).group(1)) == _lib.MAX_GROUP
    # both tables travel by value in the kernel parameters: they have to stay below the 4 KB parameter space
    assert 4 + 4 * (_lib.MAX_GROUP + 1) + _lib.MAX_GROUP * ctypes.sizeof(_lib.LinearProblem) < 4096
    assert 8 + 4 * (_lib.MAX_GROUP + 1) + _lib.MAX_GROUP * (16 + ctypes.sizeof(_lib.WgradProblem)) < 4096


def test_argument_errors_are_reported_not_crashed(lib):
    from graph_recsys_benchmark_b200 import _lib
    v = _lib.CsrView()
    rc = lib.peagnn_spmm(ctypes.byref(v), None, 64, 63, None, 64, None, None, 0, None, 0, 0, None)
    assert rc < 0 and b'peagnn_spmm' in lib.peagnn_last_error()
    with pytest.raises(RuntimeError):
        _lib.call('peagnn_linear', None, 64, None, 0, 10, 63, 64, None, 0, None, 0, 0, None, 64, None, 0, None)


def test_missing_library_fails_loudly(monkeypatch):
    from graph_recsys_benchmark_b200 import _lib
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/libpeagnn_sm100.so')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch
    from graph_recsys_benchmark_b200 import nn as pnn
    conv = pnn.PEAGCNConv(8, 4)
    with pytest.raises(RuntimeError, match='CUDA only'):
        conv(torch.randn(5, 8), torch.zeros(2, 3, dtype=torch.long))
