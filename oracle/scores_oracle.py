"""torch-CPU restatement of the reference's image metrics -- TEST INFRASTRUCTURE ONLY.

  * util/scores.py:68-86   gaussian_filter      -> ``gaussian_filter``
  * util/scores.py:88-131  ssim                 -> ``ssim``
  * util/scores.py:133-173 _ssim_per_channel    -> ``_per_channel``
  * util/scores.py:11-48   img2mse / img2psnr   -> ``mse`` / ``psnr``

Parity status: PINNED -- tests/test_oracle_vs_reference.py::test_scores_match_reference compares these bit-for-bit with the
functions imported from /root/reference/util/scores.py where that tree exists."""
import torch
import torch.nn.functional as F


def gaussian_filter(size: int, sigma: float) -> torch.Tensor:
    coords = torch.arange(size).to(dtype=torch.float32)
    coords -= (size - 1) / 2.
    g = coords ** 2
    g = (- (g.unsqueeze(0) + g.unsqueeze(1)) / (2 * sigma ** 2)).exp()
    g /= g.sum()
    return g.unsqueeze(0)


def _per_channel(x, y, kernel, data_range=1., k1=0.01, k2=0.03):
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    n = x.size(1)
    mu1 = F.conv2d(x, weight=kernel, stride=1, padding=0, groups=n)
    mu2 = F.conv2d(y, weight=kernel, stride=1, padding=0, groups=n)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(x * x, weight=kernel, stride=1, padding=0, groups=n) - mu1_sq
    sigma2_sq = F.conv2d(y * y, weight=kernel, stride=1, padding=0, groups=n) - mu2_sq
    sigma12 = F.conv2d(x * y, weight=kernel, stride=1, padding=0, groups=n) - mu1_mu2
    cs_map = (2 * sigma12 + c2) / (sigma1_sq + sigma2_sq + c2)
    ssim_map = ((2 * mu1_mu2 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    return ssim_map.mean(dim=(-1, -2)), cs_map.mean(dim=(-1, -2))


def ssim(x, y, kernel_size=11, kernel_sigma=1.5, data_range=1., reduction='mean', full=False, k1=0.01, k2=0.03):
    kernel = gaussian_filter(kernel_size, kernel_sigma).repeat(x.size(1), 1, 1, 1).to(y)
    ssim_map, cs_map = _per_channel(x, y, kernel, data_range, k1, k2)
    ssim_val, cs = ssim_map.mean(1), cs_map.mean(1)
    if reduction != 'none':
        op = {'mean': torch.mean, 'sum': torch.sum}[reduction]
        ssim_val, cs = op(ssim_val, dim=0), op(cs, dim=0)
    return (ssim_val, cs) if full else ssim_val


def mse(x, y):
    return torch.mean((x - y) ** 2)


def psnr(x, y):
    return -10. * torch.log(torch.mean((x - y) ** 2)) / torch.log(torch.Tensor([10.]))
