"""Overlay data in the reference's on-disk formats, loaded ONCE into a pool instead of per update:

* `datasets/carla/*.npy`: one uint8 `(3,84,84)` frame per file, written by `utils.py:325-327` and read B files at a time by
  `augmentations.sample_frames_from_carla_dataset` (`augmentations.py:65-76`) inside every SGSAC aux update;
* Places365 (`augmentations.py:17-62`): `ImageFolder(<dir>/places365_standard/<train|val>)` with
  `RandomResizedCrop(size) -> RandomHorizontalFlip -> ToTensor`, streamed by a DataLoader for SVEA's `random_overlay`.

Both return host arrays / tensors; the agents copy them to the device (`SGSAC.set_overlay_pool`, `SVEA.set_places_pool`) and
the overlay kernels index the pool with per-step device-drawn ids.
"""
import os

import numpy as np
import torch


def load_carla_frames(path, limit=None, size=84):
    """uint8 ndarray (N,3,size,size) from `<path>/*.npy` (sorted by file name)."""
    files = sorted(f for f in os.listdir(path) if f.endswith(".npy"))
    if limit is not None:
        files = files[:int(limit)]
    if not files:
        raise FileNotFoundError(f"no .npy frames under {path}")
    out = np.empty((len(files), 3, size, size), dtype=np.uint8)
    for i, f in enumerate(files):
        a = np.load(os.path.join(path, f))
        if a.shape != (3, size, size):
            raise ValueError(f"{f}: frame shape {a.shape}, expected {(3, size, size)} (utils.py:325-327 writes (3,84,84) uint8)")
        if a.dtype != np.uint8:
            if a.min() < 0 or a.max() > 255:
                raise ValueError(f"{f}: values outside 0..255")
            a = a.astype(np.uint8)
        out[i] = a
    return out


def load_places_pool(data_dirs, n, image_size=84, use_val=False, seed=0):
    """float32 tensor (n,3,image_size,image_size) in [0,1]: n images drawn (shuffled, with the reference's random transform,
    `augmentations.py:26-35`) from the first existing directory of `data_dirs` (`utils.load_config("datasets")` in the
    reference), looking for `places365_standard/<partition>` inside it and falling back to the directory itself."""
    import torchvision.datasets as datasets
    import torchvision.transforms as TF
    if isinstance(data_dirs, (str, os.PathLike)):
        data_dirs = [data_dirs]
    partition = "val" if use_val else "train"
    for data_dir in data_dirs:
        if not os.path.exists(data_dir):
            continue
        fp = os.path.join(data_dir, "places365_standard", partition)
        if not os.path.exists(fp):
            fp = data_dir
        ds = datasets.ImageFolder(fp, TF.Compose([TF.RandomResizedCrop(image_size), TF.RandomHorizontalFlip(), TF.ToTensor()]))
        if len(ds) == 0:
            continue
        g = torch.Generator().manual_seed(seed)
        order = torch.randperm(len(ds), generator=g).tolist()
        state = torch.random.get_rng_state()
        torch.manual_seed(seed)                      # the torchvision transforms draw from torch's global generator
        try:
            imgs = [ds[order[i % len(ds)]][0] for i in range(int(n))]
        finally:
            torch.random.set_rng_state(state)
        return torch.stack(imgs).float()
    raise FileNotFoundError("failed to find places365 data at any of the specified paths")
