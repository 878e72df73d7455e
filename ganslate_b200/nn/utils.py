"""Weight init, norm selection and LR schedule -- same behaviour as ganslate/nn/utils.py:8-99."""
from torch import nn
from torch.nn import init
from torch.optim import lr_scheduler

from . import layers


def init_net(network, conf, device):
    init_weights(network, conf.train.gan.weight_init_type, conf.train.gan.weight_init_gain)
    return network.to(device)


def init_weights(net, weight_init_type='normal', gain=0.02):
    """Fill every *Conv* / *Linear* weight in module order (ganslate/nn/utils.py:13-36): identical RNG
    consumption, so the same seed gives the same weights as the reference."""

    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, 'weight') and (classname.find('Conv') != -1 or classname.find('Linear') != -1):
            if weight_init_type == 'normal':
                init.normal_(m.weight.data, 0.0, gain)
            elif weight_init_type == 'xavier':
                init.xavier_normal_(m.weight.data, gain=gain)
            elif weight_init_type == 'kaiming':
                init.kaiming_normal_(m.weight.data, a=0, mode='fan_in')
            elif weight_init_type == 'orthogonal':
                init.orthogonal_(m.weight.data, gain=gain)
            else:
                raise NotImplementedError(f"initialization method `{weight_init_type}` is not implemented")
            if hasattr(m, 'bias') and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find('BatchNorm3d') != -1:
            init.normal_(m.weight.data, 1.0, gain)
            init.constant_(m.bias.data, 0.0)

    net.apply(init_func)


def get_conv_layer_3d(is_separable=False):
    """ganslate/nn/utils.py:39-43"""
    return layers.SeparableConv3d if is_separable else layers.Conv3d


def get_conv_transpose_layer_3d(is_separable=False):
    """ganslate/nn/utils.py:46-50"""
    return layers.SeparableConvTranspose3d if is_separable else layers.ConvTranspose3d


def get_norm_layer_2d(norm_type='instance'):
    if norm_type == 'instance':
        return layers.InstanceNorm2d
    raise NotImplementedError(f"Normalization layer `{norm_type}` not supported by the B200 path "
                              "(the reference's default and every shipped config use `instance`)")


def get_norm_layer_3d(norm_type='instance'):
    if norm_type == 'instance':
        return layers.InstanceNorm3d
    raise NotImplementedError(f"Normalization layer `{norm_type}` not supported by the B200 path")


def is_bias_before_norm(norm_type='instance'):
    if norm_type == 'instance':
        return True
    elif norm_type == 'batch':
        return False
    raise NotImplementedError(f"Normalization layer `{norm_type}` not supported")


def get_scheduler(optimizer, conf):
    """Constant LR for n_iters then linear decay to zero over n_iters_decay (ganslate/nn/utils.py:83-99)."""

    def lambda_rule(iter_idx):
        start_iter = 1
        if conf.train.checkpointing.load_iter:
            start_iter += conf.train.checkpointing.load_iter
        return 1.0 - max(0, iter_idx + start_iter - conf.train.n_iters) / float(conf.train.n_iters_decay + 1)

    return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda_rule)


def get_network_device(network):
    return next(network.parameters()).device
