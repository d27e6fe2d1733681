from . import tensorbuffer  # noqa: F401
