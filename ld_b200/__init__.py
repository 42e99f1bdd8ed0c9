"""ld_b200 -- B200-native (sm_100a) implementation of the LD (Localization Distillation) training hot
path of HikariTJU/LD behind the reference's mmdet registry surface.  Importing the package registers
`LDHead`, `GFLHead`, `KnowledgeDistillationKLDivLoss`, `IMLoss`, `QualityFocalLoss`,
`DistributionFocalLoss`, `GIoULoss`, `ATSSAssigner`, `AnchorGenerator`, `ResNet`, `FPN`, `GFL`,
`KnowledgeDistillationSingleStageDetector` under the reference's names.  All compute runs in the
C-ABI CUDA library (include/ld_b200.h); there is no CPU / PyTorch fallback."""
from .config import Config, ConfigDict, Registry, build_from_cfg  # noqa: F401
from . import registry  # noqa: F401
from . import anchors, losses, backbones, dense_heads, detectors  # noqa: F401  (register modules)
from .registry import (build_detector, build_head, build_loss, build_backbone, build_neck,  # noqa: F401
                       BACKBONES, NECKS, HEADS, LOSSES, DETECTORS, BBOX_ASSIGNERS)
from .detectors import LDTrainStep  # noqa: F401

__version__ = '0.1.0'
