"""preworld_b200: B200-native (sm_100a) camera->voxel occupancy forward path of
getterupper/PreWorld behind the reference's own registry names.

    from preworld_b200 import build_model, model_cfg
    model = build_model(model_cfg('finetune', 'r50')).cuda().eval()
    out = model(return_loss=False, img_inputs=[img_inputs], img_metas=[None])
"""
from .config import Config, ConfigDict
from .configs import model_cfg
from .plugin import (BACKBONES, DETECTORS, HEADS, MODELS, NECKS,
                     build_backbone, build_detector, build_head, build_model,
                     build_neck)

__version__ = '0.1.0'
