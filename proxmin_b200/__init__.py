from . import workloads  # noqa: F401
