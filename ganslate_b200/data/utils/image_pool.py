"""History buffer of generated images shown to the discriminators -- behaviour of
ganslate/data/utils/image_pool.py:24-60 (python `random`, returns the input itself while the pool fills).

`ImagePool` is the reference's structure (a python list of 1-image tensors, `clone` + `cat` per query).
`DeviceImagePool` makes the same decisions with the same python `random` stream -- so a run is reproducible against
the reference image by image -- but keeps the history in ONE preallocated device tensor and answers a query with a
single gather + a single scatter launch (index tensors built on the host from the decisions): no per-image clone /
unsqueeze / cat launches, and the returned batch is a fresh tensor the CUDA-graph path can copy into its static
input.  Selected with `train.device_image_pool` (default on when the images live on a CUDA device)."""
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


class DeviceImagePool:
    """Same query semantics and random stream as ImagePool; storage (pool_size, *image_shape) on the images' device."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        self.num_imgs = 0
        self.store = None

    def query(self, images):
        if self.pool_size == 0:
            return images
        images = images.detach()
        B = images.shape[0]
        if self.store is None or self.store.shape[1:] != images.shape[1:] or self.store.device != images.device:
            self.store = torch.empty((self.pool_size,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
            self.num_imgs = 0
        # the reference's decisions, image by image, with the same calls to `random`
        src_slot = [-1] * B      # >= 0: the answer for image i is the stored image of that slot (as it is BEFORE the query
        dst_slot = [-1] * B      #        modified by earlier images of this batch -- handled below); dst: where image i goes
        latest = {}              # slot -> index of the batch image written to it earlier in this query
        take_new = [-1] * B      # >= 0: the answer is batch image `take_new[i]` (an image of this same batch)
        for i in range(B):
            if self.num_imgs < self.pool_size:
                dst_slot[i] = self.num_imgs
                latest[self.num_imgs] = i
                self.num_imgs += 1
                take_new[i] = i
                continue
            if random.uniform(0, 1) > 0.5:
                idx = random.randint(0, self.pool_size - 1)
                if idx in latest:
                    take_new[i] = latest[idx]     # the slot was overwritten by an earlier image of this batch
                else:
                    src_slot[i] = idx
                dst_slot[i] = idx
                latest[idx] = i
            else:
                take_new[i] = i
        out = images.clone()
        dev = images.device
        rows = [i for i in range(B) if src_slot[i] >= 0]
        if rows:
            out[torch.tensor(rows, device=dev)] = self.store[torch.tensor([src_slot[i] for i in rows], device=dev)]
        moved = [i for i in range(B) if take_new[i] >= 0 and take_new[i] != i]
        if moved:
            out[torch.tensor(moved, device=dev)] = images[torch.tensor([take_new[i] for i in moved], device=dev)]
        if latest:
            slots = list(latest.keys())
            self.store[torch.tensor(slots, device=dev)] = images[torch.tensor([latest[k] for k in slots], device=dev)]
        return out


def make_image_pool(conf):
    """`train.device_image_pool` (default True): DeviceImagePool when training on a CUDA device, else the reference's
    list-based ImagePool."""
    size = conf.train.gan.pool_size
    use_dev = bool(conf.train.get("device_image_pool", True)) and bool(conf.train.get("cuda", True)) and torch.cuda.is_available()
    return DeviceImagePool(size) if use_dev else ImagePool(size)
