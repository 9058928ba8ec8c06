from .build import make_model  # noqa: F401
from .iodine import IODINE  # noqa: F401
