"""SingleVarianceNetwork — mirror of models/modules/variance_network.py:5-11 (NeuS inv_s scalar)."""
import torch
import torch.nn as nn


class SingleVarianceNetwork(nn.Module):
    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(float(init_val))))

    def forward(self, x):
        # ones([n,1]) * exp(10 * variance): a single scalar, trivially cheap -> plain torch
        return torch.ones([len(x), 1], dtype=x.dtype, device=x.device) * torch.exp(self.variance * 10.0)

    def inv_s(self):
        """clip(exp(10*variance), 1e-6, 1e6) as used at implicit_surface.py:126 (quirk Q9)."""
        return torch.exp(self.variance.detach() * 10.0).clip(1e-6, 1e6)
