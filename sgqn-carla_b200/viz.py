"""Evaluation-time image grids of SGSAC.log_tensorboard (sgsac.py:104-135 -> rl_utils.py:85-107).  Off the update path:
plain torch indexing on the handful of frames the eval hook logs (4 observations every eval episode)."""
import torch


def make_grid(images, nrow, padding=2):
    """(N,C,H,W) -> (C, rows*(H+pad)+pad, nrow*(W+pad)+pad) image mosaic with a zero border (what torchvision.utils.make_grid
    returns for these arguments; single-channel images are repeated to 3 channels like it does)."""
    if images.size(1) == 1:
        images = images.expand(-1, 3, -1, -1)
    n, c, h, w = images.shape
    cols = min(nrow, n)
    rows = (n + cols - 1) // cols
    grid = images.new_zeros(c, rows * (h + padding) + padding, cols * (w + padding) + padding)
    for k in range(n):
        y, x = (k // cols) * (h + padding) + padding, (k % cols) * (w + padding) + padding
        grid[:, y:y + h, x:x + w] = images[k]
    return grid


def make_obs_grid(obs, n=4):
    """rl_utils.py:85-91: the three RGB frames of the first n stacks, one stack per row, scaled to [0,1]."""
    frames = torch.cat([obs[i, j:j + 3].unsqueeze(0) for i in range(n) for j in range(0, 9, 3)], 0)
    return make_grid(frames, nrow=3) / 255.0


def make_obs_grad_grid(obs_grad, n=4):
    """rl_utils.py:98-107: per frame max over its 3 channels, normalised by its own maximum, everything at or below the
    frame's 0.97 quantile zeroed."""
    sample = []
    for i in range(n):
        for j in range(0, 9, 3):
            a = obs_grad[i, j:j + 3].max(dim=0)[0]
            sample.append(a[None, None] / a.max())
    sample = torch.cat(sample, 0)
    q = torch.quantile(sample.flatten(1), 0.97, 1)
    sample = torch.where(sample <= q[:, None, None, None], torch.zeros_like(sample), sample)
    return make_grid(sample, nrow=3)


class NullWriter(object):
    """Stands in for SummaryWriter when tensorboard is not installed: accepts the calls, keeps the last images for callers
    that want them."""

    def __init__(self, log_dir=None):
        self.log_dir, self.images, self.scalars = log_dir, {}, {}

    def add_image(self, tag, img, global_step=None, **kw):
        self.images[tag] = (global_step, img.detach().cpu())

    def add_scalar(self, tag, value, global_step=None, **kw):
        self.scalars[tag] = (global_step, float(value))

    def flush(self):
        pass

    def close(self):
        pass


def make_writer(log_dir):
    """SummaryWriter(log_dir) as in sgsac.py:41-48, or the no-op writer when tensorboard is unavailable."""
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(log_dir)
    except Exception:
        return NullWriter(log_dir)
