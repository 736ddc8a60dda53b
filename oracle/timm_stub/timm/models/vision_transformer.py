"""Restatement of the two timm 1.0.9 symbols the reference imports
(`from timm.models.vision_transformer import Mlp, DropPath`, model_v1/model/HTR_VT.py:4).

TEST INFRASTRUCTURE ONLY (oracle). Semantics follow timm 1.0.9's published behaviour:
  * Mlp: fc1 -> act() -> Dropout(drop) -> Identity norm -> fc2 -> Dropout(drop);
    attribute names fc1/act/drop1/norm/fc2/drop2 (state_dict keys `mlp.fc1.*`, `mlp.fc2.*`).
  * DropPath: per-sample Bernoulli(keep) mask scaled by 1/keep in training, identity otherwise.
"""
import torch
import torch.nn as nn


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None,
                 act_layer=nn.GELU, norm_layer=None, bias=True, drop=0., use_conv=False):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
        self.act = act_layer()
        self.drop1 = nn.Dropout(drop)
        self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
        self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
        self.drop2 = nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0., scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask
