"""TEST INFRASTRUCTURE — a deterministic stand-in for the torchsparse cost-volume regularisation network
(SparseCostRegNetList, reg_network.py:87-106), which cannot run here (torchsparse 2.1.0 is un-vendored).  The same
function is plugged into the UNMODIFIED reference's SuRF.build_volumes (oracle/make_golden.py: case_build_volumes) and
into surf_b200's, so that the orchestration around it is compared on identical inputs."""
import torch
import torch.nn as nn


class StandinReg(nn.Module):          # (an nn.Module so that it can replace the reference's child module)
    def __init__(self, d_in=(8, 16, 16, 16), d_base=(8, 8, 8, 8), d_out=(8, 8, 8, 8), seed=77):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.A = [torch.randn(i, o, generator=g) * 0.7 for i, o in zip(d_in, d_out)]
        self.B = [torch.randn(i, b, generator=g) * 0.7 for i, b in zip(d_in, d_base)]

    def dense(self, feats, stage):
        """feats (n, d_in[stage]) -> (out (n, d_out), mid (n, d_base)); computed on the CPU in fp32 on both sides so that
        the two pipelines see bit-identical regulariser outputs for identical inputs."""
        f = feats.detach().float().cpu()
        out = torch.tanh(f @ self.A[stage])
        mid = torch.tanh(f @ self.B[stage])
        return out.to(feats.device), mid.to(feats.device)

    # the reference calls reg_network(sparse_tensor, stage) (surf.py:115); surf_b200 calls (feats, coords, stage)
    def forward(self, a, b, c=None):
        if c is None:
            return self.dense(a.F, b)
        return self.dense(a, c)
