"""ctypes binding of csrc/libenflow_b200.so (the C ABI declared in include/enflow_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails this raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libenflow_b200.so')

_lib = None

vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
f64, u64, box3 = C.c_double, C.c_uint64, C.POINTER(C.c_double)


class Dims(C.Structure):
    """enflow_dims_t"""
    _fields_ = [('B', i32), ('N', i32), ('nf', i32), ('L', i32), ('E_cap', i32), ('max_n', i32),
                ('dt', f32), ('coords_weight', f32), ('mode', i32), ('fc', i32)]


# name -> (restype, argtypes); mirrors include/enflow_b200.h one to one
SIGNATURES = {
    'enflow_last_error': (C.c_char_p, []),
    'enflow_version': (i32, []),
    'enflow_hidden': (i32, []),
    'enflow_launch_count': (C.c_longlong, [i32]),
    'enflow_timing_enable': (i32, [i32]),
    'enflow_timing_kinds': (i32, []),
    'enflow_timing_read': (i32, [C.POINTER(f32), C.POINTER(i32)]),
    'enflow_adam_step': (i32, [vp, vp, vp, vp, i64, vp, f32, vp, f32, f32, f32, vp]),
    'enflow_param_layout': (i64, [i32, i32, C.POINTER(i64), C.POINTER(i64)]),
    'enflow_lj_prior_workspace_doubles': (i64, [i32]),
    'enflow_lj_prior_forces': (i32, [vp, i32, box3, f64, f64, vp, vp, vp, vp]),
    'enflow_lj_prior_minimize': (i32, [vp, i32, box3, f64, f64, i32, f64, f64, vp, vp]),
    'enflow_lj_prior_velocities': (i32, [vp, i32, f64, u64, vp]),
    'enflow_lj_prior_run': (i32, [vp, vp, i32, box3, f64, f64, f64, f64, f64, i32, u64, u64, vp, vp, vp]),
    'enflow_edges_workspace_ints': (i64, [i32]),
    'enflow_build_edges': (i32, [vp, vp, i32, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
    'enflow_fc_check': (i32, [vp, vp, vp, vp, i32, vp, vp]),
    'enflow_fc_build': (i32, [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'enflow_build_col_perm': (i32, [vp, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp]),
    'enflow_segment_sum128': (i32, [vp, vp, vp, i32, i32, i32, vp, vp]),
    'enflow_segment_sum3': (i32, [vp, vp, vp, i32, i32, i32, f32, i32, vp, vp]),
    'enflow_pack_floats': (i64, [i32]),
    'enflow_pack_layer': (i32, [vp, i32, vp, vp]),
    'enflow_node_pre_fwd': (i32, [vp, i32, i32, vp, vp, vp, vp, vp]),
    'enflow_edge_fwd': (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
    'enflow_tc_pack_bytes': (i64, []),
    'enflow_tc_pack_layer': (i32, [vp, i32, vp, vp]),
    'enflow_edge_fwd_tc': (i32, [i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
    'enflow_run_rows': (i64, [i32, i32]),
    'enflow_run_scratch_ints': (i64, [i32]),
    'enflow_run_index': (i32, [vp, i32, vp, vp, vp]),
    'enflow_run_sum128': (i32, [vp, vp, vp, i32, i32, vp, vp]),
    'enflow_node_post_fwd': (i32, [vp, vp, i32, i32, vp, vp, vp, vp, vp]),
    'enflow_coupling_fwd': (i32, [vp] * 9 + [i32, i32, f32] + [vp] * 6),
    'enflow_coupling_bwd': (i32, [vp, vp, vp, i32, i32, f32] + [vp] * 8),
    'enflow_coupling_inv_pre': (i32, [vp, vp, vp, i32, i32, f32, vp, vp, vp]),
    'enflow_coupling_inv_post': (i32, [vp, vp, vp, vp, i32, i32, f32, vp, vp, vp, vp]),
    'enflow_argmax_fwd': (i32, [vp, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp]),
    'enflow_nll_slices': (i32, [i32]),
    'enflow_nll_fwd': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp]),
    'enflow_nll_bwd': (i32, [vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, vp, vp, vp, vp, vp, vp, vp]),
    'enflow_flow_workspace_bytes': (sz, [C.POINTER(Dims), i32]),
    'enflow_flow_forward': (i32, [C.POINTER(Dims)] + [vp] * 10 + [sz, i32] + [vp] * 8),
    'enflow_flow_backward': (i32, [C.POINTER(Dims)] + [vp] * 7 + [sz] + [vp] * 7),
    'enflow_flow_reverse': (i32, [C.POINTER(Dims)] + [vp] * 9 + [sz, i32] + [vp] * 3),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(or `make -C enflow_b200/csrc`). enflow_b200 has no CPU fallback.')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(rc):
    if rc != 0:
        raise RuntimeError('enflow_b200: ' + lib().enflow_last_error().decode())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('enflow_b200 runs on CUDA tensors only (no CPU fallback); got a CPU tensor')


def f32c(t):
    return t.detach().to(torch.float32).contiguous()


def param_layout(nf, L):
    n = 15 * L + 4
    offs = (i64 * n)()
    cnts = (i64 * n)()
    total = lib().enflow_param_layout(nf, L, offs, cnts)
    if total < 0:
        raise RuntimeError('enflow_b200: ' + lib().enflow_last_error().decode())
    return int(total), list(offs), list(cnts)


TIMING_KINDS = ['edges', 'node_pre', 'edge_fwd', 'run_sum', 'seg_cols', 'seg_rows', 'segment_sum3', 'node_post', 'coupling_fwd',
                'coupling_bwd', 'coupling_inv', 'edge_geom', 'edge_bwd', 'edge_reduce', 'node_post_bwd', 'node_pre_bwd',
                'col_perm', 'argmax', 'nll']


def timing_read():
    """{family: (total_ms, groups)} collected since enflow_timing_enable(1)."""
    n = lib().enflow_timing_kinds()
    ms = (f32 * n)()
    cnt = (i32 * n)()
    check(lib().enflow_timing_read(ms, cnt))
    return {TIMING_KINDS[k]: (float(ms[k]), int(cnt[k])) for k in range(n)}


MODES = {'fp32': 0, 'fp32_tc': 1, 'bf16': 2}
