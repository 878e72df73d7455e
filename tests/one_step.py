"""One eager CycleGAN iteration after a warm-up step (target command for `ncu` launch lists)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200.presets import cyclegan_resnet2d
from ganslate_b200.utils.builders import build_gan
from oracle import torch_oracle as O

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
model = build_gan(cyclegan_resnet2d(batch_size=batch))
a, b = O.synthetic_batch(batch, 3, 256, seed=1)
a, b = a.cuda(), b.cuda()
for _ in range(steps):
    model.set_input({"A": a, "B": b})
    model.optimize_parameters()
torch.cuda.synchronize()
print("done", {k: float(v) for k, v in model.losses.items() if v is not None})
