from . import inits  # noqa: F401
from . import conv  # noqa: F401
