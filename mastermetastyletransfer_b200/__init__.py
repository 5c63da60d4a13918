"""mastermetastyletransfer_b200 -- B200-native (sm_100a) stylization hot path of Master: Meta Style Transformer.

Drop-in mirrors of the reference's codes/{full_model,style_transformer,decoder,loss}.py modules whose
forward passes run hand-written CUDA kernels behind the C ABI in include/mst_b200.h.
"""
from .decoder import Decoder  # noqa: F401
from .loss import VGG19_custom, custom_loss  # noqa: F401
from .full_model import MasterStyleTransferModel, SwinEncoderB200  # noqa: F401
from .style_transformer import (ShiftedWindowAttention, ShiftedWindowAttention_for_decoder_last_MHA,  # noqa: F401
                                StyleDecoder, StyleEncoder, StyleSwinTransformerBlock, StyleTransformer)

__version__ = "0.1.0"
