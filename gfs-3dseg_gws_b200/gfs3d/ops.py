"""Thin torch-facing wrappers over the C ABI: allocate outputs as torch tensors, pass raw device pointers and the
current CUDA stream.  No arithmetic happens here."""
from typing import Optional

import torch

from ._lib import check, lib

ACT_NONE, ACT_LRELU02, ACT_RELU = 0, 1, 2


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gfs3d ops need CUDA tensors: the hot path has no CPU fallback")


def act_rows(M: int) -> int:
    return (M + 127) // 128 * 128


def new_act(M: int, kblocks: int, device) -> torch.Tensor:
    """zeroed bf16 activation matrix in the tiled SWIZZLE_128B layout: [M/128][kblocks] tiles of 16 KiB"""
    return torch.zeros(act_rows(M) // 128, kblocks, 128 * 64, dtype=torch.bfloat16, device=device)


def act_to_dense(act: torch.Tensor, M: int) -> torch.Tensor:
    """(debug / tests) undo the tiled swizzled layout -> (M, kblocks*64) bf16 row-major"""
    mt, kb, _ = act.shape
    t = act.view(mt, kb, 128, 8, 8)
    r = torch.arange(128, device=act.device).view(128, 1)
    q = torch.arange(8, device=act.device).view(1, 8)
    src = (q ^ (r & 7)).view(1, 1, 128, 8, 1).expand(mt, kb, 128, 8, 8)
    dense = torch.gather(t, 3, src)                       # chunk q of row r lives at position q ^ (r & 7)
    return dense.permute(0, 2, 1, 3, 4).reshape(mt * 128, kb * 64)[:M]


def knn(x: torch.Tensor, k: int, return_dist: bool = False):
    """x: (B, C, N) fp32 view with unit point stride and channel stride N -> idx (B, N, k) int32"""
    _need_cuda(x)
    B, C, N = x.shape
    assert x.dtype == torch.float32 and x.stride(2) == 1 and x.stride(1) == N, "x must be channel-major"
    sq = torch.empty(B, N, dtype=torch.float32, device=x.device)
    idx = torch.empty(B, N, k, dtype=torch.int32, device=x.device)
    dist = torch.empty(B, N, k, dtype=torch.float32, device=x.device) if return_dist else None
    check(lib().gfs_knn_f32(_ptr(x), x.stride(0), B, C, N, k, _ptr(sq), _ptr(idx), _ptr(dist), _stream()), "gfs_knn_f32")
    return (idx, dist) if return_dist else idx


def pointwise(x: torch.Tensor, wt: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """x (B,C,N) cm fp32, wt (C,O) fp32 -> (B*N, O) fp32"""
    _need_cuda(x, wt, bias)
    B, C, N = x.shape
    assert x.stride(2) == 1 and x.stride(1) == N and wt.is_contiguous() and wt.shape[0] == C
    O = wt.shape[1]
    out = torch.empty(B * N, O, dtype=torch.float32, device=x.device)
    check(lib().gfs_pointwise_f32(_ptr(x), x.stride(0), B, C, N, _ptr(wt), _ptr(bias), O, _ptr(out), _stream()),
          "gfs_pointwise_f32")
    return out


def pack_weight(w: torch.Tensor, row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """w (R, K) fp32 -> bf16 tiles [ceil(K/64)][R x 64], K-major SWIZZLE_128B, optional per-row scale folded in"""
    _need_cuda(w, row_scale)
    w = w.contiguous().float()
    R, K = w.shape
    kb = (K + 63) // 64
    out = torch.empty(kb, R * 64, dtype=torch.bfloat16, device=w.device)
    rs = None if row_scale is None else row_scale.contiguous().float()
    check(lib().gfs_pack_weight_bf16(_ptr(w), _ptr(rs), R, K, _ptr(out), _stream()), "gfs_pack_weight_bf16")
    return out


def cm_to_act(x: torch.Tensor, act: torch.Tensor, kb0: int):
    _need_cuda(x, act)
    B, C, N = x.shape
    assert x.stride(2) == 1 and x.stride(1) == N
    check(lib().gfs_cm_to_act(_ptr(x), x.stride(0), B, C, N, _ptr(act), act.shape[1], kb0, _stream()), "gfs_cm_to_act")


def edgeconv(pq, idx, w2_packed, shift2, B, N, k, y_cm=None, y_act=None, y_act_kb=0, y_act2=None, y_act2_kb=0,
             argmax=None):
    """fused EdgeConv given the graph; writes into the provided destinations"""
    _need_cuda(pq, idx, w2_packed, shift2)
    if y_cm is not None:
        assert y_cm.stride(2) == 1 and y_cm.stride(1) == N and y_cm.shape[1] == 64
    check(lib().gfs_edgeconv_fwd(
        _ptr(pq), _ptr(idx), _ptr(w2_packed), _ptr(shift2), B, N, k,
        _ptr(y_cm), 0 if y_cm is None else y_cm.stride(0),
        _ptr(y_act), 0 if y_act is None else y_act.shape[1], y_act_kb,
        _ptr(y_act2), 0 if y_act2 is None else y_act2.shape[1], y_act2_kb,
        _ptr(argmax), _stream()), "gfs_edgeconv_fwd")
