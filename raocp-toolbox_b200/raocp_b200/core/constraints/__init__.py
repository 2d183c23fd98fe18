from .base_constraint import *  # noqa: F401,F403
from .no_constraint import *  # noqa: F401,F403
from .rectangle import *  # noqa: F401,F403
