"""clipself_b200 — B200-native CLIPSelf distillation step behind the reference's open_clip / training API."""
from .configs import list_models  # noqa: F401
from .factory import create_model, create_model_and_transforms, get_cast_dtype  # noqa: F401
from .loss import ClipLoss  # noqa: F401
from .model import CustomCLIP, EVAVisionTransformer  # noqa: F401

__version__ = "0.1.0"
