from .configuration import (PhysConfig, LorenzConfig, RosslerConfig, CylinderConfig, AutoPhysConfig,
                            CONFIG_MAPPING)

__all__ = ["PhysConfig", "LorenzConfig", "RosslerConfig", "CylinderConfig", "AutoPhysConfig", "CONFIG_MAPPING"]
