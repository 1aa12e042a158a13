from .utils import distl_projection, soft_update, td_target  # noqa: F401
