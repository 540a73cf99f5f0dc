"""Seeded synthetic inputs of the benchmark workload: per-lead standardised fp32 signals and float
multi-hot labels with ~3 of 71 classes set, as the reference's dataset feeds them (`dataset.py:93`,
`ptb_dataset.py:68-77`); seed 77 is the reference's own (`config.json:1050`)."""
import torch


def synthetic_batch(batch_size, num_channels=12, length=2500, num_class=71, seed=77):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch_size, num_channels, length, generator=g, dtype=torch.float32)
    y = (torch.rand(batch_size, num_class, generator=g) < 3.0 / 71.0).float()
    return x, y
