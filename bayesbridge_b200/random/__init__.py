from .random import BasicRandom, DevicePolyaGamma, DeviceTiltedStable
