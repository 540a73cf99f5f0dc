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


def _upload(model, hyper, values):
    """asynchronous upload of the 64-byte hyper-parameter block through a ring of pinned slots (a copy from pageable
    memory would make the host wait for everything queued on the stream, i.e. for the previous step)"""
    ring = getattr(model, '_opt_ring', None)
    if ring is None:
        ring = model._opt_ring = _lib.PinnedRing(len(values), torch.float32)
    ring.upload(hyper, values)


def clip_grad_norm_(model, max_norm, norm_type=2.0, error_if_nonfinite=False):
    """`nn.utils.clip_grad_norm_` over all parameters of an `EcgVit` (train.py:281); returns the total norm."""
    if float(norm_type) != 2.0:
        raise NotImplementedError('only the 2-norm is implemented')
    lib = _lib.load()
    g = _flat_grads(model)
    hyper, stats = _scratch(model)
    _upload(model, hyper, _lib.adamw_hyper(0.0, 0.9, 0.999, 1e-8, 0.0, 1, float(max_norm), 1.0))
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.ecgvit_grad_sumsq(g.data_ptr(), _lib.F32, g.numel(), hyper.data_ptr(), stats.data_ptr(), st), 'grad_sumsq')
    _lib.check(lib.ecgvit_grad_scale_by_clip(g.data_ptr(), g.numel(), hyper.data_ptr(), stats.data_ptr(), st),
               'grad_scale_by_clip')
    total_norm = stats[2].clone()   # a device scalar, like torch's; only error_if_nonfinite reads it on the host
    if error_if_nonfinite and not math.isfinite(float(total_norm)):
        raise RuntimeError(
            f'The total norm of order {float(norm_type)} for gradients from `parameters` is non-finite, so it cannot '
            f'be clipped. To disable this error and scale the gradients by the non-finite norm anyway, set '
            f'`error_if_nonfinite=False`')
    return total_norm


class FusedAdamW(torch.optim.Optimizer):
    """AdamW over the model's flat parameter buffer; one group, decay on everything (train.py:242-244).

    `state_dict()` / `load_state_dict()` carry the moments and the step count (as flat tensors under `state['flat']`),
    so the usual `torch.save(optimizer.state_dict())` checkpoint resumes exactly.  Parameters whose `.grad` is None are
    left untouched (value and moments), as torch.optim.AdamW leaves them."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.model = model
        super().__init__(model.parameters(), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._step = 0
        self._m = self._v = None

    def _ensure_moments(self, g):
        n = g.numel()
        if self._m is None or self._m.numel() != n or self._m.device != g.device:
            old = (self._m, self._v)
            self._m, self._v = torch.zeros_like(g), torch.zeros_like(g)
            if old[0] is not None and old[0].numel() == n:   # moved device / loaded on the CPU before the first step
                self._m.copy_(old[0])
                self._v.copy_(old[1])

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        model = self.model
        model._prepare(next(model.parameters()).device)
        # parameters without a gradient keep their value and moments (torch skips them); they are rare (frozen tensors)
        frozen = [k for (k, _), p in zip(model._layout.items(), model._param_list()) if p.grad is None]
        g = _flat_grads(model)
        n = g.numel()
        self._ensure_moments(g)
        keep = []
        for k in frozen:
            o, cnt, _, _ = model._layout[k]
            keep.append((o, cnt, model._flat_p[o:o + cnt].clone(), self._m[o:o + cnt].clone(), self._v[o:o + cnt].clone()))
        grp = self.param_groups[0]
        self._step += 1
        hyper, stats = _scratch(model)
        b1, b2 = grp['betas']
        # max_norm 0: clipping is a separate call in the reference loop
        _upload(model, hyper, _lib.adamw_hyper(grp['lr'], b1, b2, grp['eps'], grp['weight_decay'], self._step, 0.0, 1.0))
        stats.zero_()
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.load().ecgvit_adamw_step(model._flat_p.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                                 g.data_ptr(), _lib.F32, _lib.ptr(model._shadow), n, hyper.data_ptr(),
                                                 stats.data_ptr(), 0, st), 'adamw_step')
        for o, cnt, p0, m0, v0 in keep:
            model._flat_p[o:o + cnt].copy_(p0)
            self._m[o:o + cnt].copy_(m0)
            self._v[o:o + cnt].copy_(v0)
        if keep:
            model.sync_shadow(force=True)
        return loss

    # ---- checkpointing: torch.optim.Optimizer.state is empty here (the moments are flat), so carry them explicitly ----
    def state_dict(self):
        sd = super().state_dict()
        sd['flat'] = {'step': self._step,
                      'exp_avg': None if self._m is None else self._m.detach().clone(),
                      'exp_avg_sq': None if self._v is None else self._v.detach().clone()}
        return sd

    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        flat = state_dict.pop('flat', None)
        super().load_state_dict(state_dict)
        if flat is not None:
            self._step = int(flat['step'])
            if flat['exp_avg'] is not None:
                dev = next(self.model.parameters()).device
                self._m = flat['exp_avg'].detach().to(device=dev, dtype=torch.float32).clone()
                self._v = flat['exp_avg_sq'].detach().to(device=dev, dtype=torch.float32).clone()
            else:
                self._m = self._v = None
