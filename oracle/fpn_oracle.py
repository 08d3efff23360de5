"""TEST INFRASTRUCTURE — CPU restatement of the reference's FeatureNetwork.forward (models/modules/feature_network.py:
126-178) with plain torch functional ops; the checker of surf_b200/modules/feature_network.py.  Only tests/ may import it.
Pinned: bit-equal to the unmodified reference module on tests/golden/fpn.npz (oracle/make_golden.py case_fpn)."""
import torch
import torch.nn.functional as F


def _cir(x, w, stride):                  # Conv2d (:6-25): conv (no bias) -> InstanceNorm2d -> ReLU
    return F.relu(F.instance_norm(F.conv2d(x, w, None, stride, 1), eps=1e-5))


def _dir(x, w):                          # Deconv2d (:56-75)
    return F.relu(F.instance_norm(F.conv_transpose2d(x, w, None, 2, 1, 1), eps=1e-5))


def feature_network_forward(sd, x, prefix=""):
    """sd: state dict of a FeatureNetwork; x (nv, 3, H, W) -> list coarse -> fine (:151-178)."""
    n = sum(1 for k in sd if k.startswith(prefix + "out_layers."))
    e_outs = []
    for i in range(n):
        x = _cir(x, sd[prefix + "encoder_layers.%d.0.conv.weight" % i], 2 if i > 0 else 1)
        x = _cir(x, sd[prefix + "encoder_layers.%d.1.conv.weight" % i], 1)
        e_outs.append(x)
    d_outs = [e_outs[-1]]
    for i in range(n - 2, -1, -1):
        d_outs.append(_dir(d_outs[-1], sd[prefix + "decoder_layers.%d.conv.weight" % i]) + e_outs[i])
    d_outs = d_outs[::-1]
    outs = [F.conv2d(d_outs[i], sd[prefix + "out_layers.%d.weight" % i], None, 1, 1) for i in range(n)]
    return outs[::-1]
