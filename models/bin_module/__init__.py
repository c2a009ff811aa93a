from . import binarized_modules  # noqa: F401
