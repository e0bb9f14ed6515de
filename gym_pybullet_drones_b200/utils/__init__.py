from .enums import ActionType, DroneModel, ImageType, ObservationType, Physics  # noqa: F401
from .Logger import Logger  # noqa: F401
