from . import policies  # noqa: F401
