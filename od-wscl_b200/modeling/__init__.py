from . import registry  # noqa: F401
from .detector import GeneralizedRCNN, build_detection_model  # noqa: F401
