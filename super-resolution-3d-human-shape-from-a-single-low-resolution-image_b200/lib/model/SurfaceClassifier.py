"""Drop-in for the reference's ``lib/model/SurfaceClassifier.py``: the parameter container of the
occupancy MLP (state-dict keys ``conv{i}.weight`` [Cout,Cin,1], ``conv{i}.bias``) plus a torch
``forward`` for the variants the fused kernels do not cover.  At inference the parameters are
handed to ``surs_set_weights`` and evaluated by csrc/query_*.cu."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class SurfaceClassifier(nn.Module):
    def __init__(self, filter_channels, num_views=1, no_residual=True, res_layers=(), last_op=None):
        super().__init__()
        self.filter_channels = list(filter_channels)
        self.num_views = num_views
        self.no_residual = no_residual
        self.res_layers = list(res_layers)
        self.last_op = last_op
        self.n_layers = len(self.filter_channels) - 1
        for l in range(self.n_layers):
            cin = self.filter_channels[l]
            if not no_residual and l in self.res_layers:
                cin += self.filter_channels[0]          # skip input appended after y (reference :63-64)
            self.add_module("conv%d" % l, nn.Conv1d(cin, self.filter_channels[l + 1], 1))

    def layers(self):
        return [getattr(self, "conv%d" % l) for l in range(self.n_layers)]

    def forward(self, feature):
        """reference lib/model/SurfaceClassifier.py:45-81 (incl. the multi-view mean pooling :70-76)."""
        y, skip = feature, feature
        for i, conv in enumerate(self.layers()):
            if not self.no_residual and i in self.res_layers:
                y = torch.cat([y, skip], 1)
            y = conv(y)
            if i != self.n_layers - 1:
                y = F.leaky_relu(y)
            if self.num_views > 1 and i == self.n_layers // 2:
                y = y.view(-1, self.num_views, y.shape[1], y.shape[2]).mean(dim=1)
                skip = feature.view(-1, self.num_views, feature.shape[1], feature.shape[2]).mean(dim=1)
        return self.last_op(y) if self.last_op else y
