"""Config presets for the BASELINE.json workloads (shapes from the reference's shipped YAMLs)."""
from ganslate_b200.configs.utils import init_config


def cyclegan_resnet2d(batch_size=1, lambda_identity=0.0, n_residual_blocks=9, **train_overrides):
    """projects/horse2zebra/experiments/default.yaml:28-49 -- Resnet2D-9 + PatchGAN2D(n_layers 3), lambda 10,
    proportion_ssim 0, lsgan, lr 2e-4.  (`multi_stream=True` as a train override: the two cycle chains on two streams.)"""
    conf = {
        "mode": "train",
        "train": {
            "batch_size": batch_size, "cuda": True, "mixed_precision": False, "n_iters": 200000, "n_iters_decay": 0,
            "gan": {
                "_target_": "ganslate_b200.nn.gans.unpaired.CycleGAN",
                "pool_size": 50,
                "generator": {"_target_": "ganslate_b200.nn.generators.Resnet2D",
                              "n_residual_blocks": n_residual_blocks, "in_out_channels": {"AB": [3, 3]}},
                "discriminator": {"_target_": "ganslate_b200.nn.discriminators.PatchGAN2D", "n_layers": 3,
                                  "in_channels": {"B": 3}},
                "optimizer": {"lambda_AB": 10.0, "lambda_BA": 10.0, "lambda_identity": lambda_identity,
                              "proportion_ssim": 0.0, "lr_D": 0.0002, "lr_G": 0.0002},
            },
        },
    }
    conf["train"].update(train_overrides)
    return init_config(conf)


def pix2pix_resnet2d(batch_size=8, lambda_pix2pix=30.0, n_residual_blocks=9, n_layers=4, **train_overrides):
    """projects/cityscapes_label2photo/experiments/pix2pix.yaml shapes with the Resnet2D generator:
    D sees cat[A, B] (6 channels), PatchGAN2D n_layers 4, lambda 30."""
    conf = {
        "mode": "train",
        "train": {
            "batch_size": batch_size, "cuda": True, "mixed_precision": False, "n_iters": 200000, "n_iters_decay": 0,
            "gan": {
                "_target_": "ganslate_b200.nn.gans.paired.Pix2PixConditionalGAN",
                "generator": {"_target_": "ganslate_b200.nn.generators.Resnet2D",
                              "n_residual_blocks": n_residual_blocks, "in_out_channels": {"AB": [3, 3]}},
                "discriminator": {"_target_": "ganslate_b200.nn.discriminators.PatchGAN2D", "n_layers": n_layers,
                                  "in_channels": {"B": 6}},
                "optimizer": {"lambda_pix2pix": lambda_pix2pix, "lr_D": 0.0002, "lr_G": 0.0002},
            },
        },
    }
    conf["train"].update(train_overrides)
    return init_config(conf)


def pix2pix_unet2d(batch_size=8, lambda_pix2pix=30.0, num_downs=7, ngf=128, use_dropout=True, n_layers=4,
                   **train_overrides):
    """projects/cityscapes_label2photo/experiments/pix2pix.yaml: Unet2D(num_downs 7, ngf 128, dropout) + PatchGAN2D
    (n_layers 4) on cat[A, B], lambda 30."""
    conf = pix2pix_resnet2d(batch_size, lambda_pix2pix, 9, n_layers, **train_overrides)
    conf.train.gan.generator = init_config({"g": {"_target_": "ganslate_b200.nn.generators.Unet2D", "num_downs": num_downs,
                                                  "ngf": ngf, "use_dropout": use_dropout,
                                                  "in_out_channels": {"AB": [3, 3]}}}).g
    return conf


def cut_resnet2d(batch_size=1, n_residual_blocks=9, **train_overrides):
    """CUT defaults of ganslate/nn/gans/unpaired/cut.py:16-40 on Resnet2D + PatchGAN2D."""
    conf = {
        "mode": "train",
        "train": {
            "batch_size": batch_size, "cuda": True, "mixed_precision": False, "n_iters": 200000, "n_iters_decay": 0,
            "gan": {
                "_target_": "ganslate_b200.nn.gans.unpaired.CUT",
                "generator": {"_target_": "ganslate_b200.nn.generators.Resnet2D",
                              "n_residual_blocks": n_residual_blocks, "in_out_channels": {"AB": [3, 3]}},
                "discriminator": {"_target_": "ganslate_b200.nn.discriminators.PatchGAN2D", "n_layers": 3,
                                  "in_channels": {"B": 3}},
                "optimizer": {"lr_D": 0.0002, "lr_G": 0.0002},
            },
        },
    }
    conf["train"].update(train_overrides)
    return init_config(conf)


def _vnet3d_conf(target, channels, use_inverse, batch_size, first_layer_channels, down_blocks, up_blocks, ndf, n_layers,
                 train_overrides, generator=None, use_memory_saving=False):
    conf = {
        "mode": "train",
        "train": {
            "batch_size": batch_size, "cuda": True, "mixed_precision": False, "n_iters": 200000, "n_iters_decay": 0,
            "gan": {
                "_target_": target,
                "pool_size": 50,
                "generator": {"_target_": "ganslate_b200.nn.generators.Vnet3D", "use_memory_saving": use_memory_saving,
                              "use_inverse": use_inverse, "first_layer_channels": first_layer_channels,
                              "down_blocks": list(down_blocks), "up_blocks": list(up_blocks),
                              "in_out_channels": {"AB": [channels, channels]}},
                "discriminator": {"_target_": "ganslate_b200.nn.discriminators.PatchGAN3D", "n_layers": n_layers,
                                  "ndf": ndf, "kernel_size": [4, 4, 4], "in_channels": {"B": channels}},
                "optimizer": {"lambda_AB": 10.0, "lambda_BA": 10.0, "lambda_identity": 0.0, "proportion_ssim": 0.0,
                              "lr_D": 0.0002, "lr_G": 0.0002},
            },
        },
    }
    if generator is not None:
        conf["train"]["gan"]["generator"] = generator
    conf["train"].update(train_overrides)
    return init_config(conf)


def cyclegan_vnet3d(channels=1, batch_size=1, first_layer_channels=16, down_blocks=(1, 2, 3, 2), up_blocks=(2, 2, 1, 1),
                    ndf=64, n_layers=3, **train_overrides):
    """BASELINE config 4: CycleGAN with two Vnet3D generators (invertible layers disabled) + PatchGAN3D on
    1x32x256x256 CBCT -> CT shaped patches (projects/maastro_lung_proton_cbct_to_ct)."""
    return _vnet3d_conf("ganslate_b200.nn.gans.unpaired.CycleGAN", channels, False, batch_size, first_layer_channels,
                        down_blocks, up_blocks, ndf, n_layers, train_overrides)


def revgan_vnet3d(channels=4, batch_size=1, first_layer_channels=16, down_blocks=(1, 2, 3, 2), up_blocks=(2, 2, 1, 1),
                  ndf=64, n_layers=3, use_memory_saving=True, **train_overrides):
    """BASELINE config 5: RevGAN, one partially invertible Vnet3D (use_inverse, inverse-recompute backward =
    use_memory_saving, the Vnet3D default -- vnet3d.py:36) + PatchGAN3D on 4x128^3 patches."""
    return _vnet3d_conf("ganslate_b200.nn.gans.unpaired.RevGAN", channels, True, batch_size, first_layer_channels,
                        down_blocks, up_blocks, ndf, n_layers, train_overrides, use_memory_saving=use_memory_saving)


def revgan_piresnet3d(channels=1, batch_size=1, depth=5, first_layer_channels=32, ndf=64, n_layers=2,
                      use_memory_saving=True, **train_overrides):
    """The shipped BraTS RevGAN experiment (projects/brats_mri_sequence_translation/experiments/revgan.yaml:25-39):
    RevGAN with one partially invertible Piresnet3D (depth 5, Piresnet3DConfig default first_layer_channels 32) and
    PatchGAN3D(n_layers 2) on 1x32x176x176 patches."""
    gen = {"_target_": "ganslate_b200.nn.generators.Piresnet3D", "use_memory_saving": use_memory_saving, "use_inverse": True,
           "first_layer_channels": first_layer_channels, "depth": depth, "in_out_channels": {"AB": [channels, channels]}}
    return _vnet3d_conf("ganslate_b200.nn.gans.unpaired.RevGAN", channels, True, batch_size, first_layer_channels, (), (),
                        ndf, n_layers, train_overrides, generator=gen)
