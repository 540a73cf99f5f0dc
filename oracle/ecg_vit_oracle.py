"""
ORACLE (test infrastructure, NOT product code) -- travelling CPU restatement of the reference hot path.

Restates, in plain PyTorch on CPU,
  * the `EcgVit` wrapper                       /root/reference/ecg_transformer/models/ecg_vit.py:95-149
  * its named sizes                            /root/reference/ecg_transformer/models/ecg_vit.py:56-92
  * one training step of `MyTrainer.train`     /root/reference/ecg_transformer/models/train.py:241-283
  * the LR schedules the trainer picks         /root/reference/ecg_transformer/models/train.py:245-252
    (third party: transformers==4.17.0 `get_{cosine,constant}_schedule_with_warmup`, restated below)
over the restated third-party `vit-pytorch==0.33.2` ViT (`oracle/vit_restated.py`).

PARITY UNPINNED: the reference holds no golden vectors for this path.  In the build
container this file is validated against the reference's verbatim wrapper (`oracle/ref_shim.py`,
`tests/test_oracle.py::test_oracle_matches_reference_wrapper`); on the GPU box it is validated against
the committed fixtures `tests/golden/*.npz` which were generated THROUGH the reference wrapper
(`tests/golden/make_golden.py`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import math
from collections import namedtuple

import torch
from torch import nn

from oracle.vit_restated import ViT

ModelOutput = namedtuple('ModelOutput', ['loss', 'logits'])  # util/models.py:3

# ecg_vit.py:64-91 -- (hidden, layers, heads, mlp)
NAMED_SIZES = {
    'debug': (64, 4, 4, 256),
    'tiny': (256, 4, 4, 1024),
    'small': (512, 8, 8, 2048),
    'base': (768, 12, 12, 3072),
    'large': (1024, 24, 16, 4096),
}


class OracleConfig:
    """Field names and defaults of `EcgVitConfig` (ecg_vit.py:29-54)."""

    def __init__(self, max_signal_length=2560, patch_size=64, num_channels=12, hidden_size=512,
                 num_hidden_layers=8, num_attention_heads=8, intermediate_size=2048,
                 hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, num_class=71, per_lead_tokens=False):
        self.max_signal_length = max_signal_length
        self.patch_size = patch_size
        self.num_channels = num_channels
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.num_class = num_class
        # BASELINE.json configs[3] ("per-lead tokens"): not constructible through the reference wrapper;
        # its oracle is vit_pytorch's ViT(image_size=(C, L), patch_size=(1, P), channels=1) on [B, 1, C, L]
        self.per_lead_tokens = per_lead_tokens
        self.size = None

    @classmethod
    def from_defined(cls, model_name):
        name, size = model_name.rsplit('-', 1)
        assert name == 'ecg-vit' and size in NAMED_SIZES
        c = cls()
        c.size = size
        c.hidden_size, c.num_hidden_layers, c.num_attention_heads, c.intermediate_size = NAMED_SIZES[size]
        return c


class OracleEcgVit(nn.Module):
    """`EcgVit` (ecg_vit.py:95-149): ViT over [B, C, 1, L], BCE-with-logits on the CLS head."""

    def __init__(self, num_class=71, config=None, loss_reduction='mean'):
        super().__init__()
        config = config if config is not None else OracleConfig()
        d, h = config.hidden_size, config.num_attention_heads
        assert d % h == 0  # ecg_vit.py:99
        self.config = config
        self.per_lead = bool(getattr(config, 'per_lead_tokens', False))
        self.vit = ViT(  # ecg_vit.py:102-116; note the dropout wiring quirk (:113-114)
            image_size=(config.num_channels if self.per_lead else 1, config.max_signal_length),
            patch_size=(1, config.patch_size), num_classes=num_class,
            dim=d, depth=config.num_hidden_layers, heads=h, mlp_dim=config.intermediate_size, pool='cls',
            channels=1 if self.per_lead else config.num_channels, dim_head=d // h,
            dropout=config.hidden_dropout_prob, emb_dropout=config.attention_probs_dropout_prob)
        self.loss_reduction = loss_reduction

    def forward(self, sample_values, labels=None):
        if self.per_lead:
            logits = self.vit(sample_values.unsqueeze(1))   # [B, 1, C, L]: leads are image rows, one channel
        else:
            logits = self.vit(sample_values.unsqueeze(-2))  # ecg_vit.py:141
        loss = None
        if labels is not None:
            loss = nn.functional.binary_cross_entropy_with_logits(logits, labels, reduction=self.loss_reduction)
        return ModelOutput(loss=loss, logits=logits)


def lr_lambda(schedule, step, n_warmup, n_total):
    """transformers 4.17 `get_{constant,cosine}_schedule_with_warmup` multipliers (train.py:245-252)."""
    if step < n_warmup:
        return float(step) / float(max(1, n_warmup))
    if schedule == 'constant':
        return 1.0
    progress = float(step - n_warmup) / float(max(1, n_total - n_warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * 0.5 * 2.0 * progress)))


class OracleTrainer:
    """The step of `MyTrainer.train` (train.py:241-283) without logging / data loading."""

    def __init__(self, model, learning_rate=3e-4, weight_decay=1e-2, schedule='constant', n_warmup=0,
                 n_step=1000, max_grad_norm=1.0):
        self.model = model
        # one group, decay on everything (train.py:242-244)
        self.optimizer = torch.optim.AdamW(model.parameters(), lr=learning_rate, weight_decay=weight_decay)
        self.scheduler = torch.optim.lr_scheduler.LambdaLR(
            self.optimizer, lambda s: lr_lambda(schedule, s, n_warmup, n_step))
        self.max_grad_norm = max_grad_norm

    def step(self, sample_values, labels):
        self.optimizer.zero_grad()                                              # train.py:271
        out = self.model(sample_values=sample_values, labels=labels)            # train.py:275
        out.loss.backward()                                                     # train.py:280
        total_norm = nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=self.max_grad_norm,
                                              error_if_nonfinite=True)          # train.py:281
        self.optimizer.step()                                                   # train.py:282
        self.scheduler.step()                                                   # train.py:283
        return out.loss.detach(), out.logits.detach(), total_norm.detach()


def synthetic_batch(batch_size, num_channels=12, length=2500, num_class=71, seed=77):
    """seed 77 (reference `config.json:1050`), standardised fp32 leads, multi-hot labels."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch_size, num_channels, length, generator=g, dtype=torch.float32)
    y = (torch.rand(batch_size, num_class, generator=g) < 3.0 / 71.0).float()
    return x, y


def patch_matrix(x, patch_size):
    """Integer-exact statement of the Rearrange: A[b*n+w, t*C+c] = x[b, c, w*P+t]."""
    b, c, l = x.shape
    n = l // patch_size
    return x.reshape(b, c, n, patch_size).permute(0, 2, 3, 1).reshape(b * n, patch_size * c).contiguous()
