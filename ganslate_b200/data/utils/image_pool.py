"""History buffer of generated images shown to the discriminators -- behaviour of
ganslate/data/utils/image_pool.py:24-60 (python `random`, returns the input itself while the pool fills)."""
import random

import torch


class ImagePool:

    def __init__(self, pool_size):
        self.pool_size = pool_size
        if self.pool_size > 0:
            self.num_imgs = 0
            self.images = []

    def query(self, images):
        if self.pool_size == 0:
            return images
        picked = []
        for image in images:
            image = torch.unsqueeze(image.data, 0)
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(image)
                picked.append(image)
                continue
            if random.uniform(0, 1) > 0.5:
                idx = random.randint(0, self.pool_size - 1)
                old = self.images[idx].clone()
                self.images[idx] = image
                picked.append(old)
            else:
                picked.append(image)
        return torch.cat(picked, 0)
