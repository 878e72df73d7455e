"""Config presets for the BASELINE.json workloads (shapes from the reference's shipped YAMLs)."""
from ganslate_b200.configs.utils import init_config


def cyclegan_resnet2d(batch_size=1, lambda_identity=0.0, n_residual_blocks=9, **train_overrides):
    """projects/horse2zebra/experiments/default.yaml:28-49 -- Resnet2D-9 + PatchGAN2D(n_layers 3), lambda 10,
    proportion_ssim 0, lsgan, lr 2e-4."""
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
