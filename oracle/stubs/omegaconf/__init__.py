"""Import-only stand-in so that /root/reference's nn modules import without OmegaConf (test infrastructure)."""
MISSING = "???"


def II(x):
    return "${" + x + "}"


class DictConfig(dict):
    pass


class OmegaConf:
    pass


class dictconfig:
    DictConfig = DictConfig
