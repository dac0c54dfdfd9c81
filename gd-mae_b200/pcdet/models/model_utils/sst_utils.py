"""Reference-format views of the window bookkeeping (pcdet/models/model_utils/sst_utils.py:6-181)
derived from ops.WindowTable.  The hot path never calls these (it works on flat tokens); they
exist so that callers / tests that want the reference's tensors - batch_win_inds, coors_in_win,
per-level flat2win indices, padded window tensors - get them with the reference's semantics."""
import torch

from .... import ops as _ops

DROP_LEVEL_TOKENS = {0: 16, 1: 32, 2: 64}  # DROP_INFO of every GD-MAE config (gd_mae_ssl.yaml:63-75)


def get_window_coors(table):
    """-> (batch_win_inds (N,) int64, coors_in_win (N,3) int64 [z,y,x]) as sst_utils.get_window_coors:6-47."""
    win = table.win_of_token.long() * 2  # max_num_win_z == 2 and the z window index is 0
    pos = table.pos_of_token.long()
    return win, torch.stack([torch.zeros_like(pos), pos // 8, pos % 8], dim=-1)


def get_flat2win_inds_v2(table):
    """sst_utils.py:68-104: {level: (flat2window_inds, (positions,))} + 'voxel_drop_level'."""
    out = {}
    lvl = table.level.long()
    rank = table.lvl_rank.long()[table.win_of_token.long()]
    for dl, mt in DROP_LEVEL_TOKENS.items():
        m = lvl == dl
        if not bool(m.any()):
            continue
        out[dl] = (rank[m] * mt + table.inner.long()[m], torch.where(m))
    out['voxel_drop_level'] = lvl
    out['batching_info'] = {dl: {'max_tokens': mt} for dl, mt in DROP_LEVEL_TOKENS.items()}
    return out


def flat2window_v2(feat, inds_dict):
    """sst_utils.py:107-148."""
    out = {}
    for dl in [k for k in inds_dict if not isinstance(k, str)]:
        inds, (pos,) = inds_dict[dl]
        mt = DROP_LEVEL_TOKENS[dl]
        nwin = int(torch.div(inds, mt, rounding_mode='floor').max()) + 1
        buf = feat.new_zeros((nwin * mt,) + tuple(feat.shape[1:]))
        buf[inds] = feat[pos]
        out[dl] = buf.reshape((nwin, mt) + tuple(feat.shape[1:]))
    return out


def window2flat_v2(feat_3d_dict, inds_dict):
    """sst_utils.py:151-181."""
    n = inds_dict['voxel_drop_level'].shape[0]
    first = next(iter(feat_3d_dict.values()))
    out = first.new_zeros((n, first.shape[-1]))
    for dl, f in feat_3d_dict.items():
        inds, (pos,) = inds_dict[dl]
        out[pos] = f.reshape(-1, f.shape[-1])[inds]
    return out


def get_inner_win_inds(win_inds):
    return _ops.ingroup_inds(win_inds.contiguous())
