from nfb200.core.dtype import *  # noqa: F401,F403
from nfb200.core.dtype import DType, get_default_dtype, set_default_dtype
