"""ctypes binding of ``libpeagnn_sm100.so`` (C ABI declared in ``include/peagnn.h``).

The shared library is the product; this module only loads it, declares the argument types and
turns non-zero return codes into ``RuntimeError(peagnn_last_error())`` - the reference's error
convention is Python exceptions.  There is no fallback: if the library is
missing, importing any compute path of this package raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpeagnn_sm100.so')


class CsrView(C.Structure):
    """Mirror of ``peagnn_csr_t``."""
    _fields_ = [
        ('rowptr', C.c_void_p), ('col', C.c_void_p),
        ('nrows', C.c_int32), ('row_offset', C.c_int32),
        ('heavy_threshold', C.c_int32), ('n_heavy', C.c_int32),
        ('heavy_rows', C.c_void_p), ('heavy_chunk_ptr', C.c_void_p),
        ('n_chunks', C.c_int32),
        ('chunk_row', C.c_void_p), ('chunk_begin', C.c_void_p), ('chunk_end', C.c_void_p),
        ('partial', C.c_void_p),
        ('nnz', C.c_int64),
        ('explicit_self_loops', C.c_int32), ('sparse_filter', C.c_int32),
        ('active_rows', C.c_void_p), ('active_cols', C.c_void_p),
    ]


class LinearProblem(C.Structure):
    """Mirror of ``peagnn_linear_problem_t``."""
    _fields_ = [('X', C.c_void_p), ('ldx', C.c_int64), ('n', C.c_int64), ('W', C.c_void_p), ('bias', C.c_void_p),
                ('Y', C.c_void_p), ('ldy', C.c_int64), ('out_mask', C.c_void_p), ('ldom', C.c_int64)]


class WgradProblem(C.Structure):
    """Mirror of ``peagnn_wgrad_problem_t``."""
    _fields_ = [('X', C.c_void_p), ('ldx', C.c_int64), ('dY', C.c_void_p), ('ldd', C.c_int64), ('n', C.c_int64),
                ('dW', C.c_void_p), ('db', C.c_void_p)]


MAX_GROUP = 32      # PEAGNN_MAX_GROUP

_P, _I32, _I64, _INT, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_int, C.c_float, C.c_size_t
_U64 = C.c_uint64
_G = C.POINTER(CsrView)

# name -> (restype, argtypes); every name here must be declared in include/peagnn.h
SIGNATURES = {
    'peagnn_version': (_INT, []),
    'peagnn_last_error': (C.c_char_p, []),
    'peagnn_launch_count': (C.c_ulonglong, []),
    'peagnn_csr_workspace_bytes': (_SZ, [_I64, _I32]),
    'peagnn_csr_build': (_INT, [_P, _P, _I64, _I32, _INT, _P, _P, _P, _P, _SZ, _P]),
    'peagnn_csr_filter_workspace_bytes': (_SZ, [_I64]),
    'peagnn_csr_filter': (_INT, [_G, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    'peagnn_degree_scale': (_INT, [_P, _I32, _F, _F, _INT, _P, _P]),
    'peagnn_partial_floats': (_SZ, [_I32, _I32, _I32]),
    'peagnn_spmm': (_INT, [_G, _P, _I64, _I32, _P, _I64, _P, _P, _INT, _P, _INT, _INT, _P]),
    'peagnn_spmm_filtered': (_INT, [_G, _P, _I64, _I32, _P, _I64, _P, _P, _INT, _P, _INT, _INT, _P, _P, _P]),
    'peagnn_spmm_proj': (_INT, [_G, _P, _I64, _P, _I64, _P, _P, _INT, _P, _I32, _P, _P, _P, _P, _P, _P, _I64, _INT, _P]),
    'peagnn_to_bf16': (_INT, [_P, _I64, _I64, _I32, _P, _I64, _P]),
    'peagnn_spmm_bf16': (_INT, [_G, _P, _I64, _I32, _P, _I64, _P, _P, _INT, _P, _INT, _INT, _P, _P, _P]),
    'peagnn_mark_rows': (_INT, [_P, _I64, _I32, _I32, _P, _P]),
    'peagnn_gat_rowmax': (_INT, [_G, _P, _P, _I32, _F, _P, _P]),
    'peagnn_gat_aggregate': (_INT, [_G, _P, _I64, _I32, _I32, _P, _P, _F, _P, _P, _P, _I64, _P, _INT, _P]),
    'peagnn_gat_backward_dst': (_INT, [_G, _P, _I64, _I32, _I32, _P, _P, _F, _P, _P, _P, _I64, _P, _P, _I64,
                                       _P, _P, _P, _P, _P]),
    'peagnn_gat_backward_src': (_INT, [_G, _P, _P, _P, _P, _P, _I64, _I32, _I32, _P, _I64, _P, _P]),
    'peagnn_linear': (_INT, [_P, _I64, _P, _I64, _I64, _I32, _I32, _P, _INT, _P, _INT, _INT, _P, _I64, _P, _I64, _P]),
    'peagnn_linear_grouped': (_INT, [_P, _I32, _I32, _I32, _INT, _INT, _INT, _P]),
    'peagnn_wgrad_grouped_workspace_floats': (_SZ, [_I32, _I32, _I32]),
    'peagnn_linear_wgrad_grouped': (_INT, [_P, _I32, _I32, _I32, _INT, _P, _SZ, _P]),
    'peagnn_wgrad_workspace_floats': (_SZ, [_I64, _I32, _I32]),
    'peagnn_linear_wgrad': (_INT, [_P, _I64, _P, _I64, _P, _I64, _I64, _I32, _I32, _INT, _P, _P, _P, _SZ, _P]),
    'peagnn_relu_backward': (_INT, [_P, _I64, _P, _I64, _I64, _I32, _P, _I64, _P]),
    'peagnn_gat_scores': (_INT, [_P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _P]),
    'peagnn_gat_scores_backward': (_INT, [_P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _P, _I64, _INT, _P, _P,
                                          _P, _SZ, _P]),
    'peagnn_fuse_forward': (_INT, [_P, _I64, _I64, _I32, _I32, _P, _INT, _INT, _P, _I64, _P]),
    'peagnn_fuse_workspace_floats': (_SZ, [_I64, _I32, _I32]),
    'peagnn_fuse_backward': (_INT, [_P, _I64, _I64, _I32, _I32, _P, _INT, _P, _I64, _P, _I64, _P, _P, _SZ, _P]),
    'peagnn_predict': (_INT, [_P, _I64, _I32, _P, _P, _I64, _P, _P, _P, _P, _P, _P]),
    'peagnn_bpr_workspace_floats': (_SZ, [_I64, _I32]),
    'peagnn_bpr_loss': (_INT, [_P, _I64, _I32, _P, _I32, _I64, _P, _P, _P, _P, _P, _INT, _P, _I64, _P, _P, _P, _P,
                               _P, _SZ, _P]),
    'peagnn_entity_workspace_floats': (_SZ, [_I64, _I32, _INT]),
    'peagnn_entity_reg': (_INT, [_P, _I64, _I32, _P, _I64, _F, _P, _INT, _P, _I64, _P, _SZ, _P]),
    'peagnn_eval_rank': (_INT, [_P, _I64, _I32, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P, _P]),
    'peagnn_column_mean': (_INT, [_P, _I64, _I64, _I32, _P, _P, _SZ, _P]),
    'peagnn_probe_out_floats': (_SZ, []),
    'peagnn_probe_gather': (_INT, [_P, _I64, _I32, _P, _I64, _P, _P]),
    'peagnn_bpr_rows': (_INT, [_P, _I64, _P, _I64, _I32, _U64, _U64, _I32, _I64, _I64, _I64, _P, _P, _I32,
                               _P, _P, _P, _P, _P, _I32, _P, _P]),
}

_lib = None
launch_count = 0     # kernels-launching C-ABI calls made by this process (bench.py reports it)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            'libpeagnn_sm100.so not found at %s - build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`graph_recsys_benchmark_b200/csrc/build.sh`; there is no CPU fallback.' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().peagnn_last_error().decode('utf-8', 'replace')


profile = None      # when a list: every call appends (tag or entry point, algorithmic bytes, start event, end event)
_fns = {}           # entry point name -> bound ctypes function


def call(name, *args, tag=None, nbytes=0):
    """Invoke an int-returning entry point; raise on a non-zero code.  With ``profile`` set, the launch is
    bracketed by CUDA events on the current stream - ``external`` ones while a CUDA graph is being captured,
    so they become event-record nodes that every replay re-records and ``elapsed_time`` can read."""
    global launch_count
    fn = _fns.get(name)
    if fn is None:
        fn = _fns[name] = getattr(load(), name)
    if profile is not None:
        import torch
        ext = torch.cuda.is_current_stream_capturing()
        e0 = torch.cuda.Event(enable_timing=True, external=ext)
        e1 = torch.cuda.Event(enable_timing=True, external=ext)
        e0.record()
        rc = fn(*args)
        e1.record()
        profile.append((tag or name, nbytes, e0, e1))
    else:
        rc = fn(*args)
    launch_count += 1
    if rc != 0:
        raise RuntimeError('%s failed (%d): %s' % (name, rc, last_error()))


def query(name, *args):
    """Invoke a size-returning entry point."""
    return getattr(load(), name)(*args)
