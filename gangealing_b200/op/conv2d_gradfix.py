"""conv2d_gradfix -- same call surface as reference models/stylegan2/op/conv2d_gradfix.py:22-75.

On torch >= 1.9 the reference's custom-autograd branch is disabled (conv2d_gradfix.py:85-92) and both
functions fall straight through to cuDNN via torch.nn.functional; that library call is what stays here
(the convolutions remain cuDNN, they are not one of the hand-written ops).
`no_weight_gradients()` is kept as a working context manager for API parity.
"""
import contextlib

from torch.nn import functional as F

enabled = False
weight_gradients_disabled = False


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    try:
        yield
    finally:
        weight_gradients_disabled = old


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    return F.conv2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding, dilation=dilation,
                    groups=groups)


def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    return F.conv_transpose2d(input=input, weight=weight, bias=bias, stride=stride, padding=padding,
                              output_padding=output_padding, dilation=dilation, groups=groups)
