from .configuration import (PhysConfig, LorenzConfig, RosslerConfig, CylinderConfig, GrayScottConfig, AutoPhysConfig,
                            CONFIG_MAPPING)

__all__ = ["PhysConfig", "LorenzConfig", "RosslerConfig", "CylinderConfig", "GrayScottConfig", "AutoPhysConfig", "CONFIG_MAPPING"]
