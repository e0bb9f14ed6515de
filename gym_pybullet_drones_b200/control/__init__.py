from .BaseControl import BaseControl  # noqa: F401
from .DSLPIDControl import DSLPIDControl  # noqa: F401
