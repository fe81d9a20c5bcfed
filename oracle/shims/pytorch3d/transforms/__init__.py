from .rotation_conversions import *  # noqa: F401,F403
