"""Mirror of pcdet/models/backbones_2d/sst_bev_backbone.py:6-42 (SSTBEVBackbone): a stack of Conv2d(3x3, optional dilation)
+ BatchNorm2d(eps 1e-3, momentum 0.01) + ReLU blocks on the dense BEV map, with identity shortcuts on the blocks listed in
CONV_SHORTCUT whenever the shape is preserved.  Same parameter names (``conv_layer.{i}.{0,1}``).  Dense convolutions are
plain library calls (cuDNN through torch, channels-last); nothing in this module is irregular."""
import torch
import torch.nn as nn


class SSTBEVBackbone(nn.Module):
    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        channels = model_cfg.NUM_FILTER
        self.conv_shortcut = list(model_cfg.CONV_SHORTCUT)
        blocks = []
        for kw in model_cfg.CONV_KWARGS:
            kw = dict(kw)
            blocks.append(nn.Sequential(nn.Conv2d(channels, **kw, bias=False),
                                        nn.BatchNorm2d(kw['out_channels'], eps=1e-3, momentum=0.01),
                                        nn.ReLU(inplace=True)))
            channels = kw['out_channels']
        self.conv_layer = nn.ModuleList(blocks)
        self.num_bev_features = channels

    def forward(self, data_dict):
        out = data_dict['spatial_features']
        for i, block in enumerate(self.conv_layer):
            y = block(out)
            out = y + out if (y.shape == out.shape and i in self.conv_shortcut) else y
        data_dict['spatial_features_2d'] = out
        return data_dict
