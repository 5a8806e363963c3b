"""Launch wrappers for the plain tensor-core GEMMs and their precision / layout helpers
(csrc/gemm_tc.cu sed_gemm_tc, csrc/conv_wgrad_tc.cu sed_gemm_tn_tc, csrc/gru.cu helpers)."""
import torch

from . import _lib
from . import ops
from ._lib import call, ptr, stream_of

BF16, F32 = torch.bfloat16, torch.float32


def split3(x2d, which):
    """fp32 (R, K) -> bf16 (R, 3K): which=0 left operand [hi|lo|hi], which=1 right operand [hi|hi|lo]."""
    x2d = x2d.contiguous()
    r, k = x2d.shape
    out = torch.empty((r, 3 * k), dtype=BF16, device=x2d.device)
    with torch.cuda.device(x2d.device):
        call('sed_split_bf16x3', x2d.data_ptr(), r, k, which, out.data_ptr(), stream_of(x2d))
    return out


def transpose_bf16(x2d):
    """fp32 (R, C) -> bf16 (C, R)."""
    x2d = x2d.contiguous()
    r, c = x2d.shape
    out = torch.empty((c, r), dtype=BF16, device=x2d.device)
    with torch.cuda.device(x2d.device):
        call('sed_transpose_to_bf16', x2d.data_ptr(), r, c, out.data_ptr(), stream_of(x2d))
    return out


def gemm_nt(a_bf16, bw_bf16, bias=None):
    """out (M, N) fp32 = A (M, K) @ Bw (N, K)^T + bias."""
    m, k = a_bf16.shape
    n = bw_bf16.shape[0]
    assert bw_bf16.shape[1] == k and a_bf16.is_contiguous() and bw_bf16.is_contiguous()
    out = torch.empty((m, n), dtype=F32, device=a_bf16.device)
    with torch.cuda.device(a_bf16.device):
        call('sed_gemm_tc', a_bf16.data_ptr(), bw_bf16.data_ptr(), ptr(bias), out.data_ptr(), m, n, k,
             stream_of(a_bf16))
    return out


def linear_x3(x2d_f32, w_f32, bias=None):
    """x @ w^T + bias with fp32-class accuracy on the bf16 tensor cores (3-way operand split)."""
    return gemm_nt(split3(x2d_f32, 0), split3(w_f32, 1), bias)


def gemm_tn(a_bf16, b_bf16, m, n, out, a_col=0, b_col=0, accumulate=False, ldb=None):
    """out (m, n) fp32 = A[:, a_col:a_col+m]^T @ B[:, b_col:b_col+n]; A (R, lda), B (R, ldb) bf16."""
    r, lda = a_bf16.shape
    ldb = b_bf16.shape[1] if ldb is None else ldb
    assert b_bf16.shape[0] == r and a_bf16.is_contiguous() and b_bf16.is_contiguous()
    with torch.cuda.device(a_bf16.device):
        splits = _lib.lib().sed_gemm_tn_tc_splits(r, m, n)
        slabs = torch.empty((splits, m, n), dtype=F32, device=a_bf16.device)
        call('sed_gemm_tn_tc', a_bf16.data_ptr() + 2 * a_col, lda, b_bf16.data_ptr() + 2 * b_col, ldb,
             slabs.data_ptr(), r, m, n, stream_of(a_bf16))
    return ops.reduce_partials(slabs, out, accumulate=accumulate)


def colsum(x2d, out):
    r, c = x2d.shape
    with torch.cuda.device(x2d.device):
        P = _lib.lib().sed_stat_partials()
        partial = torch.empty((P, c), dtype=F32, device=x2d.device)
        call('sed_colsum_f32', x2d.data_ptr(), r, c, partial.data_ptr(), stream_of(x2d))
    return ops.reduce_partials(partial, out)
