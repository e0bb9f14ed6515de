from .enums import ActionType, DroneModel, ImageType, ObservationType, Physics  # noqa: F401
