"""Synthetic S3DIS / ScanNet-shaped inputs for tests, smoke and bench (no dataset is reachable: no network).

Blocks follow what dataloaders/loader.py:87-101 emits: ch0-2 xyz after min-subtraction (x, y in [0,1] for a 1 m block,
z in [0,3]), ch3-5 rgb in [0,1], ch6-8 XYZ normalised per block."""
import torch


def synthetic_blocks(B: int, N: int, seed: int = 1234, dup_frac: float = 0.0) -> torch.Tensor:
    out = torch.empty(B, 9, N, dtype=torch.float32)
    scale = torch.tensor([[1.0], [1.0], [3.0]])
    for b in range(B):
        g = torch.Generator().manual_seed(seed + b)
        xyz = torch.rand(3, N, generator=g) * scale
        rgb = torch.rand(3, N, generator=g)
        if dup_frac > 0:      # sampling with replacement (loader.py:66) produces exact duplicates
            nd = int(N * dup_frac)
            src = torch.randint(0, N, (nd,), generator=g)
            dst = torch.randperm(N, generator=g)[:nd]
            xyz[:, dst] = xyz[:, src]
            rgb[:, dst] = rgb[:, src]
        xyz = xyz - xyz.min(dim=1, keepdim=True).values
        out[b, 0:3] = xyz
        out[b, 3:6] = rgb
        out[b, 6:9] = xyz / xyz.max(dim=1, keepdim=True).values.clamp_min(1e-12)
    return out


def randomize_bn_(module: torch.nn.Module, seed: int = 5):
    """non-trivial BatchNorm affine / running statistics so that folding is exercised"""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            n = m.num_features
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.3 * torch.randn(n, generator=g))
                m.bias.copy_(0.2 * torch.randn(n, generator=g))
                m.running_mean.copy_(0.2 * torch.randn(n, generator=g))
                m.running_var.copy_(0.5 + 1.5 * torch.rand(n, generator=g))
    return module
