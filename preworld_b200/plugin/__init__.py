"""Host-side mirror of the reference's plugin (registry) interface for the
camera->voxel path.  Importing this package registers the modules under the
reference's ``type`` names."""
from .builder import (BACKBONES, DETECTORS, HEADS, LOSSES, MODELS, NECKS,
                      build_backbone, build_detector, build_head, build_loss,
                      build_model, build_neck)
from .image import CustomFPN, FPN_LSS, ResNet
from .swin import SwinTransformer
from .view_transformer import LSSViewTransformerBEVStereo
from .voxel_encoder import CustomResNet3D, LSSFPN3D
from .heads import NerfHead, OccHead
from .detectors import BEVStereo4DOCC, PreWorld, PreWorld4DTraj

__all__ = ['MODELS', 'BACKBONES', 'NECKS', 'HEADS', 'LOSSES', 'DETECTORS',
           'build_backbone', 'build_neck', 'build_head', 'build_loss',
           'build_detector', 'build_model', 'ResNet', 'CustomFPN',
           'SwinTransformer', 'FPN_LSS',
           'LSSViewTransformerBEVStereo', 'CustomResNet3D', 'LSSFPN3D',
           'OccHead', 'NerfHead', 'BEVStereo4DOCC', 'PreWorld',
           'PreWorld4DTraj']
