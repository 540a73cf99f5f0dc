"""
Fused training step with the semantics of `MyTrainer.train`'s inner loop
(/root/reference/ecg_transformer/models/train.py:268-283):

    zero_grad -> model(**inputs) -> loss.backward() -> clip_grad_norm_(1.0, error_if_nonfinite=True)
              -> AdamW.step() -> scheduler.step()

but with no autograd graph, no per-tensor optimizer launches and no host synchronisation: the loss stays on the
device (the reference syncs twice per step for logging, train.py:278,289), the LR schedule runs on the host and
reaches the kernels through a 64-byte device block, and the non-finite check of `clip_grad_norm_` is a device
flag polled on request.  With `world_size > 1` gradients are summed over NCCL in per-layer buckets on a side
stream while backward is still running (`parallel.py`).
"""
import collections
import math

import torch

from . import _lib



def lr_multiplier(schedule, step, n_warmup, n_total):
    """transformers' `get_constant_schedule_with_warmup` / `get_cosine_schedule_with_warmup` (train.py:245-252)."""
    if step < n_warmup:
        return float(step) / float(max(1, n_warmup))
    if schedule == 'constant':
        return 1.0
    progress = float(step - n_warmup) / float(max(1, n_total - n_warmup))
    return max(0.0, 0.5 * (1.0 + math.cos(math.pi * progress)))


def get_train_args(args=None, n_train=None):
    """defaults of the reference trainer (train.py:407-436) that matter to the step"""
    out = dict(num_train_epoch=3, train_batch_size=64, optimizer='AdamW', learning_rate=3e-4, weight_decay=1e-2,
               warmup_ratio=0.05, schedule='cosine', max_grad_norm=1.0)
    out.update(args or {})
    if out['optimizer'] not in ('AdamW',):
        raise ValueError("only optimizer='AdamW' is implemented in the fused step")
    if out['schedule'] not in ('constant', 'cosine'):
        raise ValueError(f"Unexpected schedule: expect one of ['constant', 'cosine'], got {out['schedule']!r}")
    if n_train is not None:
        out['steps_per_epoch'] = n_train // out['train_batch_size']
        out['n_step'] = out['steps_per_epoch'] * out['num_train_epoch']
    return out


class FusedTrainer:
    def __init__(self, model, learning_rate=3e-4, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8,
                 schedule='constant', n_warmup=0, n_step=1 << 30, max_grad_norm=1.0, process_group=None,
                 bucket_layers=1, use_cuda_graph=False, data_parallel=True):
        self.model = model
        self.lr, self.wd, self.betas, self.eps = learning_rate, weight_decay, betas, eps
        self.schedule, self.n_warmup, self.n_step = schedule, n_warmup, n_step
        self.max_grad_norm = max_grad_norm
        self.step_count = 0
        self.lib = _lib.load()
        self.group = process_group
        self.world = 1
        if data_parallel and (process_group is not None or
                              (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        self.bucket_layers = bucket_layers
        self.use_cuda_graph = use_cuda_graph
        self._state_ready = False
        self._graph = None
        self._hyper_ring = None
        self._reducer = None
        self.launches_per_step = None
        self._step_done = collections.deque(maxlen=2)

    # ---- lazily created device state ---------------------------------------------------------------
    def _ensure_state(self, device):
        m = self.model
        m._prepare(device)
        if self._state_ready and self.exp_avg.device == device and self.exp_avg.numel() == m._flat_p.numel():
            return
        n = m._flat_p.numel()
        self.exp_avg = torch.zeros(n, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=device, dtype=torch.float32)
        self.hyper = torch.zeros(16, device=device, dtype=torch.float32)
        self.stats = torch.zeros(_lib.STATS_FLOATS, device=device, dtype=torch.float32)
        if self.world > 1:
            from .parallel import BucketedGradReducer
            self._reducer = BucketedGradReducer(m, self.group, self.bucket_layers)
            m._engine.base_seed += 7919 * torch.distributed.get_rank(self.group)  # independent masks per replica
        self._state_ready = True

    def current_lr(self):
        return self.lr * lr_multiplier(self.schedule, self.step_count, self.n_warmup, self.n_step)

    def _upload_hyper(self):
        t = self.step_count + 1  # AdamW's own step counter starts at 1
        b1, b2 = self.betas
        vals = _lib.adamw_hyper(self.current_lr(), b1, b2, self.eps, self.wd, t,
                                self.max_grad_norm if self.max_grad_norm is not None else 0.0, 1.0 / self.world)
        if self._hyper_ring is None:
            self._hyper_ring = _lib.PinnedRing(len(vals), torch.float32)
        self._hyper_ring.upload(self.hyper, vals)  # asynchronous: the host keeps queueing steps ahead of the device

    # ---- the step ----------------------------------------------------------------------------------
    def _device_step(self, sample_values, labels):
        m, st = self.model, torch.cuda.current_stream().cuda_stream
        eng = m._engine
        side = eng.side_stream
        if side is not None:
            # the 342 MB gradient memset does not depend on the forward pass: on the side stream it becomes a parallel
            # branch of the step graph and disappears behind the (tensor-bound) forward kernels
            main = torch.cuda.current_stream()
            fork = torch.cuda.Event()
            fork.record(main)
            side.wait_event(fork)  # after the previous step's AdamW / all-reduce, which read the gradients
            with torch.cuda.stream(side):
                m._flat_g.zero_()
                zeroed = torch.cuda.Event()
                zeroed.record(side)
        loss, logits = eng.forward(sample_values, labels, m.loss_reduction)
        if side is not None:
            main.wait_event(zeroed)
        if self._reducer is not None:
            self._reducer.begin()
        eng.backward(grad_scale=1.0, zero_grads=side is None)
        if self._reducer is not None:
            self._reducer.finish()
        n = m._flat_g.numel()
        _lib.check(self.lib.ecgvit_grad_sumsq(m._flat_g.data_ptr(), n, self.hyper.data_ptr(), self.stats.data_ptr(), st),
                   'grad_sumsq')
        _lib.check(self.lib.ecgvit_adamw_step(m._flat_p.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                              m._flat_g.data_ptr(), _lib.ptr(m._shadow), n, self.hyper.data_ptr(),
                                              self.stats.data_ptr(), st), 'adamw_step')
        return loss, logits

    def step(self, sample_values, labels, time_out_spans=None):
        """One optimisation step on device-resident fp32 inputs; returns (loss, logits) device tensors (no sync).

        The returned tensors are views of workspace buffers: valid until the next step.  With an `input_pipeline` on
        the model, `sample_values` are raw records; `time_out_spans` as in `EcgVit.forward`."""
        self._ensure_state(sample_values.device)
        self._upload_hyper()
        pipe = self.model.input_pipeline
        if pipe is not None and pipe.timeout is not None and self.model.training:
            if time_out_spans is None:
                time_out_spans = pipe.draw_spans(sample_values.shape[0], pipe.padded_length(sample_values.shape[2]))
            self.model._engine.set_spans(time_out_spans)  # a persistent device buffer: the captured graph reads it
        cfg = self.model.config
        if self.model.training and (cfg.hidden_dropout_prob > 0 or cfg.attention_probs_dropout_prob > 0):
            self.model._engine.new_dropout_seed()  # a device scalar: the captured graph reads the fresh value
        if self.use_cuda_graph:
            out = self._graph_step(sample_values, labels)
        else:
            out = self._device_step(sample_values, labels)
        self.step_count += 1  # scheduler.step() (train.py:283)
        ev = torch.cuda.Event()
        ev.record()
        self._step_done.append(ev)  # lets stage() reuse an input slot only after the step that read it
        return out

    def _graph_step(self, sample_values, labels):
        lw = self.model.loss_weight
        pipe = self.model.input_pipeline
        key = (tuple(sample_values.shape), tuple(labels.shape), tuple(float(v) for v in lw) if lw else None,
               None if pipe is None else (id(pipe), pipe.pad, pipe.timeout, self.model.training))
        if self._graph is None or self._graph_key != key:
            self._static_x = torch.empty_like(sample_values)
            self._static_y = torch.empty(labels.shape, device=labels.device, dtype=torch.float32)
            self._static_x.copy_(sample_values)
            self._static_y.copy_(labels)
            # eager warm-up (sets function attributes, allocates workspaces) on a side stream, then capture;
            # parameters/optimizer state are snapshotted so the warm-up does not count as a step
            snap = (self.model._flat_p.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone())
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._device_step(self._static_x, self._static_y)
            torch.cuda.current_stream().wait_stream(s)
            self.model._flat_p.copy_(snap[0]); self.exp_avg.copy_(snap[1]); self.exp_avg_sq.copy_(snap[2])
            self.model.sync_shadow(force=True)
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_counter[0]
            with torch.cuda.graph(g):
                self._graph_out = self._device_step(self._static_x, self._static_y)
            self.launches_per_step = _lib.launch_counter[0] - n0  # kernels of ours replayed by every graph launch
            self._graph, self._graph_key = g, key
        self._static_x.copy_(sample_values, non_blocking=True)
        self._static_y.copy_(labels, non_blocking=True)
        self._graph.replay()
        return self._graph_out

    # ---- input staging ---------------------------------------------------------------------------------
    def stage(self, sample_values_host, labels_host):
        """Start the host->device copy of the NEXT batch on a side stream (pinned source recommended) and return device
        tensors whose use on the current stream is ordered after the copy.  Two slots alternate, so when calls go
        `cur = stage(); loop: step(*cur); nxt = stage(); ...; cur = nxt` the copy of batch i+1 overlaps the kernels of
        step i (the reference copies synchronously inside the step, train.py:274)."""
        dev = torch.device('cuda', torch.cuda.current_device())
        if getattr(self, '_stage_slots', None) is None or self._stage_slots[0][0].shape != sample_values_host.shape:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_slots = [(torch.empty(sample_values_host.shape, device=dev, dtype=torch.float32),
                                  torch.empty(labels_host.shape, device=dev, dtype=torch.float32),
                                  torch.cuda.Event()) for _ in range(2)]
            self._stage_next = 0
        xd, yd, copied = self._stage_slots[self._stage_next]
        self._stage_next ^= 1
        cs = self._copy_stream
        # this slot was last read by the step BEFORE the most recently enqueued one (which reads the other slot)
        if len(self._step_done) == 2:
            cs.wait_event(self._step_done[0])
        elif len(self._step_done) == 1 and self.step_count >= 2:
            cs.wait_event(self._step_done[0])
        with torch.cuda.stream(cs):
            xd.copy_(sample_values_host, non_blocking=True)
            yd.copy_(labels_host, non_blocking=True)
            copied.record(cs)
        torch.cuda.current_stream().wait_event(copied)
        return xd, yd

    # ---- host-visible results (these synchronise) ----------------------------------------------------
    def grad_norm(self):
        return float(self.stats[2])

    def check_finite(self):
        """`error_if_nonfinite=True` of clip_grad_norm_ (train.py:281): raises like torch does, one poll late"""
        s = self.stats.tolist()
        if s[1] != 0.0 or not math.isfinite(s[2]):
            raise RuntimeError(
                'The total norm of order 2.0 for gradients from `parameters` is non-finite, so it cannot be clipped. '
                '(the fused step skipped this update)')

    def state_dict(self):
        """optimizer + schedule (+ dropout stream) state for a true resume; the reference saves the model only
        (train.py:297-300), so a restarted run there silently restarts AdamW's moments and the LR schedule"""
        self._ensure_state(next(self.model.parameters()).device)
        eng = self.model._engine
        return {'step': self.step_count, 'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone(),
                'dropout_counter': eng._seed_counter, 'dropout_base_seed': eng.base_seed}

    def load_state_dict(self, sd):
        self._ensure_state(next(self.model.parameters()).device)
        self.step_count = sd['step']
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])
        eng = self.model._engine
        eng._seed_counter = sd.get('dropout_counter', eng._seed_counter)
        eng.base_seed = sd.get('dropout_base_seed', eng.base_seed)


_TRAINERS = {}


def fused_train_step(model, batch, lr=3e-4, **trainer_kwargs):
    """One `zero_grad -> forward -> backward -> clip -> AdamW -> scheduler` step (train.py:271-283) on a batch dict as
    the reference's DataLoader yields it (`sample_values`, `labels`); the `FusedTrainer` behind it is created on first
    use and kept per model.  Returns `ModelOutput(loss, logits)` (device tensors, valid until the next step)."""
    from .model import ModelOutput
    tr = _TRAINERS.get(id(model))
    if tr is None or tr.model is not model:
        tr = _TRAINERS[id(model)] = FusedTrainer(model, learning_rate=lr, **trainer_kwargs)
    tr.lr = lr
    loss, logits = tr.step(batch['sample_values'].cuda(non_blocking=True), batch['labels'].cuda(non_blocking=True))
    return ModelOutput(loss=loss, logits=logits)
