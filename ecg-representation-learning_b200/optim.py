"""
`torch.optim`-compatible front ends for the fused flat-buffer kernels, so a reference-style loop
(/root/reference/ecg_transformer/models/train.py:242-252,271-283)

    optimizer = FusedAdamW(model, lr=lr, weight_decay=dc); scheduler = get_*_schedule_with_warmup(optimizer, ...)
    optimizer.zero_grad(); loss = model(**inputs).loss; loss.backward()
    clip_grad_norm_(model, max_norm=1.0, error_if_nonfinite=True); optimizer.step(); scheduler.step()

keeps working unchanged while each of clip / step is one or two kernel launches instead of ~1.4 k.
"""
import math

import torch

from . import _lib



def _flat_grads(model):
    """make `model._flat_g` hold the current `.grad` of every parameter and return it"""
    views = model._grad_views()
    params = model._param_list()
    src, dst = [], []
    for p, v in zip(params, views):
        if p.grad is None:
            v.zero_()
        elif p.grad.data_ptr() != v.data_ptr():
            src.append(p.grad)
            dst.append(v)
    if dst:
        torch._foreach_copy_(dst, src)
        for p, v in zip(params, views):
            if p.grad is not None:
                p.grad = v
    return model._flat_g


def _scratch(model):
    if getattr(model, '_opt_scratch', None) is None or model._opt_scratch[0].device != model._flat_p.device:
        dev = model._flat_p.device
        model._opt_scratch = (torch.zeros(16, device=dev), torch.zeros(_lib.STATS_FLOATS, device=dev))
    return model._opt_scratch


def _upload(hyper, values):
    # pageable source: the driver stages the 64 bytes before returning, so the list can be reused immediately
    hyper.copy_(torch.tensor(values, dtype=torch.float32))


def clip_grad_norm_(model, max_norm, norm_type=2.0, error_if_nonfinite=False):
    """`nn.utils.clip_grad_norm_` over all parameters of an `EcgVit` (train.py:281); returns the total norm."""
    if float(norm_type) != 2.0:
        raise NotImplementedError('only the 2-norm is implemented')
    lib = _lib.load()
    g = _flat_grads(model)
    hyper, stats = _scratch(model)
    _upload(hyper, _lib.adamw_hyper(0.0, 0.9, 0.999, 1e-8, 0.0, 1, float(max_norm), 1.0))
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.ecgvit_grad_sumsq(g.data_ptr(), g.numel(), hyper.data_ptr(), stats.data_ptr(), st), 'grad_sumsq')
    _lib.check(lib.ecgvit_grad_scale_by_clip(g.data_ptr(), g.numel(), hyper.data_ptr(), stats.data_ptr(), st),
               'grad_scale_by_clip')
    total_norm = stats[2].clone()
    if error_if_nonfinite and not math.isfinite(float(total_norm)):
        raise RuntimeError(
            f'The total norm of order {float(norm_type)} for gradients from `parameters` is non-finite, so it cannot '
            f'be clipped. To disable this error and scale the gradients by the non-finite norm anyway, set '
            f'`error_if_nonfinite=False`')
    return total_norm


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over the model's flat parameter buffer; one group, decay on everything (train.py:242-244)."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.model = model
        super().__init__(model.parameters(), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0
        self._m = self._v = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model = self.model
        model._prepare(next(model.parameters()).device)
        g = _flat_grads(model)
        n = g.numel()
        if self._m is None or self._m.numel() != n or self._m.device != g.device:
            self._m, self._v = torch.zeros_like(g), torch.zeros_like(g)
        grp = self.param_groups[0]
        self._step += 1
        hyper, stats = _scratch(model)
        b1, b2 = grp['betas']
        # max_norm 0: clipping is a separate call in the reference loop
        _upload(hyper, _lib.adamw_hyper(grp['lr'], b1, b2, grp['eps'], grp['weight_decay'], self._step, 0.0, 1.0))
        stats.zero_()
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.load().ecgvit_adamw_step(model._flat_p.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                                 g.data_ptr(), _lib.ptr(model._shadow), n, hyper.data_ptr(),
                                                 stats.data_ptr(), st), 'adamw_step')
        return loss
