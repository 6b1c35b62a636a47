"""denseclip_vit_multimodal_b200 -- B200-native (sm_100a) implementation of DenseCLIP's language-guided dense-prediction
forward path, drop-in behind the reference's class names / config keys / state_dict layout
(reference: segmentation/denseclip/__init__.py:1-3)."""
from .denseclip import DenseCLIP
from .heads import HEADS, FCNHead, IdentityHead
from .models import (BACKBONES, CLIPResNet, CLIPResNetWithAttention, CLIPTextContextEncoder, CLIPTextEncoder,
                     CLIPVisionTransformer, ContextDecoder, ViTFeatureFusionNeck)
from ._lib import DclipError, build

__all__ = ["DenseCLIP", "CLIPResNet", "CLIPTextEncoder", "CLIPVisionTransformer", "CLIPResNetWithAttention",
           "CLIPTextContextEncoder", "ContextDecoder", "IdentityHead", "FCNHead", "ViTFeatureFusionNeck", "BACKBONES", "HEADS",
           "DclipError", "build"]
