from .probabilistic_unet import ProbabilisticUnet, AxisAlignedConvGaussian, Fcomb, Encoder
from .unet import Unet
from .unet_blocks import DownConvBlock, UpConvBlock
from .utils import l2_regularisation, clean_folder
