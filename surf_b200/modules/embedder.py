"""Positional-encoding bookkeeping (models/modules/embedder.py:6-51).

The encoding itself ([x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]) is computed
inside the SDF kernel; this module only reproduces the dimension arithmetic the network
constructor needs."""


def embedder_out_dim(multires, input_dims=3):
    return input_dims * (1 + 2 * multires) if multires > 0 else input_dims
