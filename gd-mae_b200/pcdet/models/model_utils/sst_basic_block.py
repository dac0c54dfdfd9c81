"""Mirror of pcdet/models/model_utils/sst_basic_block.py:8-125 (WindowAttention, EncoderLayer,
BasicShiftBlockV2) on the flat token layout.

Argument mapping to the reference: ``pos_dict`` is the (64, d) positional table of the block
(the reference's per-level padded pos tensors hold only these 64 distinct rows), ``ind_dict`` is
the ops.WindowTable of the shift (replaces flat2win_inds + voxel_drop_level), and
``key_padding_dict`` is unused (there is no padding)."""
import torch.nn as nn
import torch.nn.functional as F

from .... import fused as _fused
from .cosine_msa import CosineMultiheadAttention


class WindowAttention(nn.Module):
    def __init__(self, d_model, nhead, dropout, batch_first=False, layer_cfg=dict()):
        super().__init__()
        self.nhead = nhead
        if not layer_cfg.get('cosine', False):
            raise NotImplementedError("GD-MAE configs use cosine attention (LAYER_CFG.cosine: True)")
        self.self_attn = CosineMultiheadAttention(d_model, nhead, dropout=dropout, batch_first=False,
                                                  tau_min=layer_cfg.get('tau_min', 0.01), cosine=True,
                                                  non_shared_tau=layer_cfg.get('non_shared_tau', False))

    def forward(self, feat_2d, pos_dict, ind_dict, key_padding_dict=None):
        return self.self_attn(feat_2d, pos_dict, ind_dict)


class EncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", batch_first=False,
                 mlp_dropout=0, layer_cfg=dict()):
        super().__init__()
        assert not batch_first and mlp_dropout == 0
        self.win_attn = WindowAttention(d_model, nhead, dropout, layer_cfg=layer_cfg)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.activation = _get_activation_fn(activation)

    fused = True  # one autograd node per layer (fused.EncoderLayerFunction); False = op-by-op autograd

    def forward(self, src, pos_dict, ind_dict, key_padding_mask_dict=None):
        if self.fused and self.activation is F.gelu:
            return _fused.encoder_layer(self, src, pos_dict, ind_dict)
        src2 = self.win_attn(src, pos_dict, ind_dict, key_padding_mask_dict)
        src = self.norm1(src + src2)
        src2 = self.linear2(self.activation(self.linear1(src)))
        return self.norm2(src + src2)


class BasicShiftBlockV2(nn.Module):
    """Two encoder layers: shift-0 windows then shift-1 windows (sst_basic_block.py:87-114)."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", batch_first=False,
                 layer_cfg=dict()):
        super().__init__()
        self.encoder_list = nn.ModuleList([
            EncoderLayer(d_model, nhead, dim_feedforward, dropout, activation, batch_first, layer_cfg=layer_cfg)
            for _ in range(2)])

    def forward(self, src, pos_dict_list, ind_dict_list, key_mask_dict_list=None):
        num_shifts = len(ind_dict_list)
        assert num_shifts in (1, 2)
        output = src
        for i in range(2):
            s = i % num_shifts
            output = self.encoder_list[i](output, pos_dict_list[s], ind_dict_list[s], None)
        return output


def _get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return F.gelu
    raise RuntimeError(F"activation should be relu/gelu, not {activation}.")
