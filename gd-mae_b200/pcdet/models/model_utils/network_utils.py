"""make_fc_layers of pcdet/models/model_utils/network_utils.py:7-21 (parameter container of the VFE MLP)."""
import torch.nn as nn


def make_fc_layers(fc_cfg, input_channels, output_channels=None, linear=True, norm_fn=None):
    assert linear
    fc_layers = []
    c_in = input_channels
    for k in range(len(fc_cfg)):
        fc_layers.extend([nn.Linear(c_in, fc_cfg[k], bias=False),
                          nn.BatchNorm1d(fc_cfg[k]) if norm_fn is None else norm_fn(fc_cfg[k]), nn.ReLU()])
        c_in = fc_cfg[k]
    if output_channels is not None:
        fc_layers.append(nn.Linear(c_in, output_channels))
    return nn.Sequential(*fc_layers)
