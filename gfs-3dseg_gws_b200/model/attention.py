"""Self-attention block -- drop-in for the reference's model/attention.py:10-48 ("adjacent").

q/k/v 1x1 maps run as ONE fused gfs_linear_bf16 call (256 -> 3*64).  The N x N softmax(q^T k / sqrt(d)) v product is the
"next" row N1 of the scope table: until the flash-style tcgen05 kernel lands it is evaluated by
torch.nn.functional.scaled_dot_product_attention (a LIBRARY kernel, flagged as such in DESIGN.md) -- the N x N matrix
is still never materialised.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from gfs3d import ops


class SelfAttention(nn.Module):
    def __init__(self, in_channel, out_channel=None, attn_dropout=0.1):
        super(SelfAttention, self).__init__()
        self.in_channel = in_channel
        self.out_channel = out_channel if out_channel is not None else in_channel
        self.temperature = self.out_channel ** 0.5
        self.q_map = nn.Conv1d(in_channel, self.out_channel, 1, bias=False)
        self.k_map = nn.Conv1d(in_channel, self.out_channel, 1, bias=False)
        self.v_map = nn.Conv1d(in_channel, self.out_channel, 1, bias=False)
        self.dropout = nn.Dropout(attn_dropout)
        self._key = None
        self._wp = None

    def _prepare(self, device):
        ws = (self.q_map.weight, self.k_map.weight, self.v_map.weight)
        key = (str(device),) + tuple((id(w), w._version) for w in ws)
        if self._key != key:
            if self.out_channel % 64 != 0 or self.in_channel % 64 != 0:
                raise NotImplementedError("SelfAttention widths must be multiples of 64 in the B200 build")
            with torch.no_grad():
                w = torch.cat([t.detach().float().reshape(self.out_channel, self.in_channel) for t in ws], dim=0)
                self._wp = ops.pack_weight(w)
            self._key = key
        return self._wp

    def forward_fused(self, x_act, B, N):
        """x_act: bf16 act tiles (B*N, in_channel) -> y (B, out_channel, N) fp32 channel-major"""
        if self.training:
            raise NotImplementedError("SelfAttention training-mode forward is not built yet in the B200 path")
        wp = self._prepare(x_act.device)
        d = self.out_channel
        qkv = torch.empty(B, 3 * d, N, dtype=torch.float32, device=x_act.device)
        ops.linear(x_act, 0, self.in_channel // 64, wp, None, 3 * d, ops.ACT_NONE, B, N, y_cm=qkv)
        q, k, v = (qkv[:, i * d:(i + 1) * d, :].transpose(1, 2).to(torch.bfloat16).unsqueeze(1) for i in range(3))
        y = F.scaled_dot_product_attention(q, k, v, scale=1.0 / self.temperature)       # (B, 1, N, d)  library kernel
        return y.squeeze(1).transpose(1, 2).float().contiguous()

    def forward(self, x):
        """(B, in_channel, N) -> (B, out_channel, N)"""
        if x.dtype != torch.float32 or x.stride(2) != 1 or x.stride(1) != x.shape[2]:
            x = x.float().contiguous()
        B, C, N = x.shape
        xa = ops.new_act(B * N, C // 64, x.device)
        ops.cm_to_act(x, xa, 0)
        return self.forward_fused(xa, B, N)
