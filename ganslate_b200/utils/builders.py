"""conf -> networks / GAN recipe: the plug-in seam of ganslate/utils/builders.py:79-129."""
from collections.abc import Mapping

from ganslate_b200.nn.utils import init_net
from ganslate_b200.utils.io import import_attr


def build_gan(conf):
    return import_attr(conf.train.gan._target_)(conf)


def build_G(conf, direction, device):
    assert direction in ['AB', 'BA']
    return build_network_by_role('generator', conf, direction, device)


def build_D(conf, domain, device):
    assert domain in ['B', 'A']
    return build_network_by_role('discriminator', conf, domain, device)


def build_network_by_role(role, conf, label, device):
    """kwargs = the role's config minus `_target_`, plus norm_type and the per-direction channel counts
    (builders.py:106-126); then init_net (nn/utils.py:8-10)."""
    assert role in ['discriminator', 'generator']
    node = conf.train.gan[role]
    network_class = import_attr(node._target_)
    args = dict(node)
    args.pop("_target_")
    args["norm_type"] = conf.train.gan.norm_type
    if role == 'generator':
        ioc = args.pop('in_out_channels')
        if isinstance(ioc, Mapping):
            ioc = ioc[label]
        args["in_channels"], args["out_channels"] = ioc
    else:
        if isinstance(args["in_channels"], Mapping):
            args["in_channels"] = args["in_channels"][label]
    return init_net(network_class(**args), conf, device)
