"""ICNR initialisation (reference models/layers/initializations.py:21-38): every r*r sub-pixel kernel of an
output channel starts identical, which removes checkerboard artefacts of the sub-pixel convolution."""
import torch
import torch.nn as nn


def ICNR(tensor, upscale_factor=2, inizializer=nn.init.kaiming_normal_):
    r2 = upscale_factor ** 2
    sub = inizializer(torch.zeros([tensor.shape[0] // r2] + list(tensor.shape[1:])))
    return sub.repeat_interleave(r2, dim=0)
