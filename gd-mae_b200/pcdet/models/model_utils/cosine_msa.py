"""Mirror of pcdet/models/model_utils/cosine_msa.py:441-528 (CosineMultiheadAttention) for the
call shape of the path: self-attention with q = k = x + pos, v = x, key padding only.

Parameters keep torch's MultiheadAttention names (in_proj_weight (3d,d) rows q|k|v, in_proj_bias,
out_proj.{weight,bias}) plus ``tau`` (1,1,1) (cosine_msa.py:452-458), so checkpoints and the
optimizer quirk (parameters held directly by this module never reach Adam
because it has the child ``out_proj``) carry over unchanged.

Forward works on FLAT tokens: one packed in-projection GEMM, the positional term folded into a
64-row look-up table (pos has 64 distinct rows), then the hand-written SRA kernel."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import ops as _ops


class CosineMultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False,
                 kdim=None, vdim=None, batch_first=False, device=None, dtype=None, cosine=True, tau_min=0.01,
                 non_shared_tau=False):
        super().__init__()
        assert bias and not add_bias_kv and not add_zero_attn and kdim is None and vdim is None
        assert dropout == 0.0, "the path trains with DROPOUT 0.0 (gd_mae_ssl.yaml:85)"
        if not cosine or non_shared_tau:
            raise NotImplementedError("the B200 SRA kernel implements the shared-tau cosine attention of the GD-MAE configs")
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, embed_dim // num_heads
        self.batch_first, self.tau_min = batch_first, tau_min
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim))
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=True)
        self.tau = nn.Parameter(torch.ones(1, 1, 1))
        self.register_buffer("_v_only", torch.cat([torch.zeros(2 * embed_dim), torch.ones(embed_dim)]), persistent=False)
        self._reset_parameters()

    def _reset_parameters(self):  # nn.MultiheadAttention._reset_parameters
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, x, pos_table, table):
        """x (N, d) flat tokens; pos_table (64, d) sin/cos embedding of the in-window cells;
        table: ops.WindowTable of the shift.  Returns (N, d) after out_proj."""
        d = self.embed_dim
        qkv = F.linear(x, self.in_proj_weight, self.in_proj_bias * self._v_only)      # q, k bias live in the LUT
        lut = F.linear(pos_table, self.in_proj_weight[:2 * d], self.in_proj_bias[:2 * d])  # (64, 2d)
        o = _ops.SraAttention.apply(qkv, lut, self.tau, table, self.tau_min, self.num_heads)
        return self.out_proj(o)
