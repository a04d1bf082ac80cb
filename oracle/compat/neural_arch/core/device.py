from nfb200.core.device import *  # noqa: F401,F403
from nfb200.core.device import Device, DeviceType, get_default_device, set_default_device
