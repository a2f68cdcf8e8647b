from .evaluate import accuracy, continuity, coverage  # noqa: F401
