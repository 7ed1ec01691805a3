from . import conv_bn_act  # noqa: F401
from .conv_bn_act import ConvBnAct, ConvNormAct, BatchNormAct2d, SEModule  # noqa: F401
