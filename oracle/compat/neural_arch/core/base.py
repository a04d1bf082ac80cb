from nfb200.core.base import Module, Optimizer, Parameter  # noqa: F401
