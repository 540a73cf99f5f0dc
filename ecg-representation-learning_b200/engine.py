"""
Step engine: owns the HBM workspaces of one (batch, length) shape and issues the kernel sequence of the
ECG-ViT forward / backward / clip+AdamW through the C ABI (include/ecgvit_b200.h).

Data layout in HBM (T = bf16 in performance mode, fp32 in parity mode; M = B * (n_patch + 1) token rows):
  parameters / gradients / AdamW moments : three flat fp32 buffers, tensors at 256-byte aligned offsets
  weight shadow                          : flat bf16 copy of the parameters (GEMM operands), written by AdamW
  a_patch  [B*n_patch, P*C]  T           : gathered patches, time-major / lead-minor features
  per layer: x_in [M,d], ln1 [M,d], qkv [M,3*inner], lse [B,H,N] f32, o [M,inner], y [M,d], ln2 [M,d],
             u [M,mlp] (pre-GELU), h [M,mlp] (post-GELU), row statistics (mean, rstd) f32.  x_in / y are fp32 when the
             residual stream is fp32 (config.residual_dtype; models deeper than 12 layers).  With
             config.activation_checkpointing only x_in is kept per layer; two rotating sets hold the rest and backward
             re-runs the block forward (same dropout masks: they are functions of the step's seed)
  backward scratch, double-buffered by layer parity (the weight-gradient GEMMs of layer l read it on the side stream while
             the main stream already writes layer l-1's): dz[2], dy[2] [M,d], dln [M,d], d_o [M,inner], dqkv[2]
             [M,3*inner], du[2] [M,mlp], dzm[site][2] [M,d] (dropout-masked LayerNorm gradients)
Everything is row-major with the feature dimension contiguous, so every GEMM operand is either K-major or
MN-major for TMA without a transposed copy.

Reference semantics: forward = vit_pytorch.ViT.forward via ecg_vit.py:140-149; backward = autograd of it
(train.py:280); optimizer = train.py:281-282.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import (GemmArgs, F32, BF16, EPI_STORE, EPI_BIAS_RES, EPI_BIAS_GELU, EPI_DGELU, EPI_ATOMIC_F32,
                   EPI_BIAS_RES_F32)

LN_EPS = 1e-5  # nn.LayerNorm default


class _Workspace:
    pass


class StepEngine:
    def __init__(self, model):
        self.model = model
        self.lib = _lib.load()
        self.ws = {}
        self._cur = None
        # device scalar holding this step's dropout seed (refreshed by the host before every training forward)
        self.device = model._flat_p.device
        self._lw_key = self._lw_table = None
        self._spans = {}
        self.rng = torch.zeros(2, device=self.device, dtype=torch.int32)
        # Weight-gradient GEMMs do not feed the backward chain: issued on a side stream they become a parallel branch of
        # the step graph and fill the SMs that the chain's kernels leave idle (partial last waves, launch ramps,
        # memory-bound kernels).  ECGVIT_WGRAD_STREAM=0 keeps everything on one stream.
        self.side_stream = torch.cuda.Stream(device=model._flat_p.device) \
            if os.environ.get('ECGVIT_WGRAD_STREAM', '1') != '0' else None
        self._seed_counter = 0
        self._seed_ring = None
        self.base_seed = 0x5EED     # rank-independent (saved in trainer state); replicas add `rank_offset`
        self.rank_offset = 0
        self._fwd_id = 0            # generation of the last forward (the autograd bridge checks it at backward)
        self.last_seed = None       # (seed, counter) uploaded last

    def new_dropout_seed(self, seed=None):
        """draw the seed of the next training forward (its backward regenerates the same masks from it)"""
        self._seed_counter += 1
        if seed is None:
            seed = ((self.base_seed + self.rank_offset) * 0x9E3779B1 + self._seed_counter * 0x85EBCA6B) & 0x7FFFFFFF
        self.upload_seed(seed, self._seed_counter & 0x7FFFFFFF)
        return seed

    def upload_seed(self, seed, counter):
        if self._seed_ring is None:
            self._seed_ring = _lib.PinnedRing(2, torch.int32)
        self._seed_ring.upload(self.rng, [seed, counter])  # asynchronous: no host stall
        self.last_seed = (seed, counter)

    def _dropout_probs(self):
        """(p_embedding, p_block) as the reference wires them (ecg_vit.py:113-114); zero in eval mode"""
        m = self.model
        if not m.training:
            return 0.0, 0.0
        return float(m.config.attention_probs_dropout_prob), float(m.config.hidden_dropout_prob)

    # ---- helpers ---------------------------------------------------------------------------------
    @property
    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def _gemm(self, M, N, K, A, lda, a_k, B, ldb, b_k, epi, out, ldo, out2=None, aux=None, bias=None, split_k=1,
              drop=None):
        p, stream = drop if drop is not None else (0.0, 0)
        g = GemmArgs(M, N, K, A.data_ptr(), lda, a_k, B.data_ptr(), ldb, b_k, epi, out.data_ptr(), ldo,
                     _lib.ptr(out2), _lib.ptr(aux), _lib.ptr(bias), self.model._dtype_code, split_k,
                     stream, p, self.rng.data_ptr() if p > 0 else None)
        if _lib.profile[0] is not None:
            keep = GemmArgs.from_buffer_copy(g)  # lets bench.py re-launch exactly this call when timing the kernel
            _lib.profile_meta[0] = (M, N, K, epi, a_k, b_k, keep)
        _lib.check(self.lib.ecgvit_gemm(ctypes.byref(g), self._stream), 'gemm')

    def workspace(self, B, L):
        key = (B, L)
        w = self.ws.get(key)
        if w is not None:
            return w
        m = self.model
        c = m.config
        dev = m._flat_p.device
        T = m._act_dtype
        P, C, d, mlp = c.patch_size, c.num_channels, c.hidden_size, c.intermediate_size
        H = c.num_attention_heads
        inner = d  # dim_head = d // heads (ecg_vit.py:100)
        assert L % P == 0, f'signal length {L} must be a multiple of patch_size {P}'
        per_lead = bool(getattr(c, 'per_lead_tokens', False))
        n = L // P * (C if per_lead else 1)
        K = P if per_lead else P * C            # features of one patch
        Kp = (K + 7) // 8 * 8 if per_lead else K  # row length of the patch matrix (16-byte aligned rows for TMA)
        assert n + 1 <= m.vit.pos_embedding.shape[1], \
            f'{n} patches exceed the positional table ({m.vit.pos_embedding.shape[1] - 1})'
        N = n + 1
        M = B * N
        depth = c.num_hidden_layers
        w = _Workspace()
        w.B, w.L, w.n, w.N, w.M = B, L, n, N, M
        w.per_lead, w.K, w.Kp = per_lead, K, Kp

        def buf(*shape, dtype=T):
            return torch.empty(*shape, device=dev, dtype=dtype)

        w.a_patch = buf(B * n, Kp)
        # per-lead tokens with P % 8 != 0: zero-padded copies of the embedding weight and of its gradient
        w.embed_wpad = buf(d, Kp) if Kp != K else None
        w.embed_gpad = buf(d, Kp, dtype=torch.float32) if Kp != K else None
        w.e = buf(B * n, d)
        RT = m._res_dtype  # residual stream: fp32 for deep models in bf16 mode (config.residual_dtype)
        w.x = [buf(M, d, dtype=RT) for _ in range(depth + 1)]  # x[l] = input of block l; x[depth] = encoder output
        # Activation checkpointing (config.activation_checkpointing; BASELINE.json configs[4]): only the block inputs x[l]
        # are kept per layer; everything inside a block lives in TWO buffer sets used alternately (layer l -> set l % 2)
        # and is recomputed from x[l] in backward.  Two sets, so that the weight-gradient GEMMs of layer l (side stream)
        # can still read their operands while layer l - 1 is being recomputed into the other set.
        w.ckpt = bool(getattr(c, 'activation_checkpointing', False)) and depth > 2
        n_sets = 2 if w.ckpt else depth

        def per_layer(make):
            sets = [make() for _ in range(n_sets)]
            return [sets[l % n_sets] for l in range(depth)]

        w.ln1 = per_layer(lambda: buf(M, d))
        w.ln2 = per_layer(lambda: buf(M, d))
        w.stat1 = per_layer(lambda: buf(2, M, dtype=torch.float32))
        w.stat2 = per_layer(lambda: buf(2, M, dtype=torch.float32))
        w.qkv = per_layer(lambda: buf(M, 3 * inner))
        w.lse = per_layer(lambda: buf(B, H, N, dtype=torch.float32))
        w.o = per_layer(lambda: buf(M, inner))
        w.y = per_layer(lambda: buf(M, d, dtype=RT))
        w.u = per_layer(lambda: buf(M, mlp))
        w.h = per_layer(lambda: buf(M, mlp))
        # head
        n_class = m.num_class
        w.xn = buf(B, d, dtype=torch.float32)
        w.hstat = buf(2, B, dtype=torch.float32)
        w.logits = buf(B, n_class, dtype=torch.float32)
        w.loss = buf(1, dtype=torch.float32)
        w.loss_none = buf(B, n_class, dtype=torch.float32)
        w.head_scratch = buf(B * d + B * n_class, dtype=torch.float32)
        w.labels = buf(B, n_class, dtype=torch.float32)
        # backward scratch
        # Every scratch buffer that a side-stream weight-gradient GEMM reads exists TWICE, indexed by layer parity: the main
        # chain of layer l - 1 then never has to wait for the weight gradients of layer l (it waits for those of layer
        # l + 1, a whole layer old), so the side stream is free to lag and fill whatever the chain leaves idle
        w.dz = [buf(M, d), buf(M, d)]          # gradient w.r.t. a block's output (input of its backward)
        w.dy = [buf(M, d), buf(M, d)]          # gradient w.r.t. the block's middle residual sum y
        w.dln = buf(M, d)
        w.d_o = buf(M, inner)
        w.dqkv = [buf(M, 3 * inner), buf(M, 3 * inner)]
        w.du = [buf(M, mlp), buf(M, mlp)]
        w.de = buf(B * n, d)
        # dropout-masked copies of the residual-stream gradients (only used when p > 0): one per site of a block
        w.dzm = [[buf(M, d), buf(M, d)], [buf(M, d), buf(M, d)]]   # [site 0 = after net[3], 1 = after to_out][parity]
        # two partial-row scratch buffers (one per LayerNorm of a block): their fold runs off the critical chain
        w.ln_scratch = [buf(int(self.lib.ecgvit_layernorm_bwd_scratch_floats(d)), dtype=torch.float32) for _ in range(2)]
        n_attn = int(self.lib.ecgvit_attention_bwd_scratch_floats(B, N, H, d // H, m._dtype_code))
        w.attn_scratch = buf(n_attn, dtype=torch.float32) if n_attn > 0 else None
        self.ws[key] = w
        return w

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, sample_values, labels=None, reduction='mean', record_attention=False):
        """sample_values: fp32 cuda [B, C, L]; returns (loss | None, logits) as views of workspace buffers.
        record_attention: also materialise every layer's softmax probabilities (slow path for `Recorder`); they are
        left in `self.recorded_attention`, fp32 [B, layers, heads, N, N]."""
        m, lib, st = self.model, self.lib, self._stream
        c = m.config
        dt = m._dtype_code
        rdt = m._res_code                                        # dtype code of calls whose operand is the residual stream
        epi_res = EPI_BIAS_RES_F32 if m._res_f32 else EPI_BIAS_RES
        x = sample_values
        assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 3, 'sample_values must be fp32 cuda [B, C, L]'
        assert x.shape[1] == c.num_channels
        if x.stride(2) != 1 or x.stride(0) != x.shape[1] * x.stride(1):
            x = x.contiguous()
        B, C, L_in = x.shape
        pipe = m.input_pipeline
        L = L_in if pipe is None else pipe.padded_length(L_in)  # TimeEndPad happens inside the gather
        w = self.workspace(B, L)
        self._cur = w
        self._fwd_id += 1
        w.fwd_id = self._fwd_id   # a later forward through the same workspace invalidates this one's backward
        P, d, mlp, H = c.patch_size, c.hidden_size, c.intermediate_size, c.num_attention_heads
        inner, dh = d, d // H
        n, N, M = w.n, w.N, w.M
        wt = m._weights()  # GEMM operand views (bf16 shadow or fp32 master)
        pf = m._params_f32()  # fp32 master views (biases, LayerNorm, cls, pos, head)

        mean = std = spans = None
        if pipe is not None:
            # raw records in: Normalize -> TimeEndPad -> TimeOut (training only) fused into the gather
            mean, std = pipe.device_stats(x.device)
            spans = self.span_buffer(B) if (pipe.timeout is not None and m.training) else None
        K, Kp = w.K, w.Kp
        if w.per_lead:
            _lib.check(lib.ecgvit_patchify_leads(x.data_ptr(), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(spans),
                                                 w.a_patch.data_ptr(), B, C, x.stride(1), L_in, L // P, P, Kp, dt, st),
                       'patchify_leads')
        elif pipe is None:
            _lib.check(lib.ecgvit_patchify(x.data_ptr(), w.a_patch.data_ptr(), B, C, x.stride(1), n, P, dt, st),
                       'patchify')
        else:
            _lib.check(lib.ecgvit_patchify_transform(x.data_ptr(), _lib.ptr(mean), _lib.ptr(std), _lib.ptr(spans),
                                                     w.a_patch.data_ptr(), B, C, x.stride(1), L_in, n, P, dt, st),
                       'patchify_transform')
        hook = m._before_layer_forward
        if hook is not None:
            hook(-1)   # embedding weights, cls, pos
        # e = a_patch @ We^T + be
        w_embed = wt['embed.w']
        if Kp != K:
            _lib.check(lib.ecgvit_pad_cols(w_embed.data_ptr(), w.embed_wpad.data_ptr(), d, K, Kp, dt, st), 'pad_cols')
            w_embed = w.embed_wpad
        self._gemm(B * n, d, Kp, w.a_patch, Kp, 1, w_embed, Kp, 1, EPI_STORE, w.e, d, bias=pf['embed.b'])
        p_emb, p_blk = self._dropout_probs()
        w.p_emb, w.p_blk = p_emb, p_blk
        seed_ptr = self.rng.data_ptr()
        _lib.check(lib.ecgvit_embed_assemble(w.e.data_ptr(), pf['cls'].data_ptr(), pf['pos'].data_ptr(),
                                             w.x[0].data_ptr(), B, n, d, p_emb, 0, seed_ptr if p_emb > 0 else None,
                                             rdt, st), 'embed_assemble')
        for l in range(c.num_hidden_layers):
            if hook is not None:
                hook(l)
            self._block_forward(l, w, record_attention=record_attention)
        if hook is not None:
            hook(c.num_hidden_layers)   # head
        red = _lib.REDUCTION[reduction]
        loss_buf = None
        if labels is not None:
            assert labels.shape == (B, m.num_class)
            w.labels.copy_(labels, non_blocking=True)  # also casts (the reference passes float multi-hot)
            loss_buf = w.loss_none if reduction == 'none' else w.loss
        _lib.check(lib.ecgvit_head_fwd(
            w.x[c.num_hidden_layers].data_ptr(), pf['head.ln.w'].data_ptr(), pf['head.ln.b'].data_ptr(),
            pf['head.w'].data_ptr(), pf['head.b'].data_ptr(), w.labels.data_ptr() if labels is not None else None,
            *self._loss_weight_table(), w.xn.data_ptr(), w.hstat[0].data_ptr(), w.hstat[1].data_ptr(), w.logits.data_ptr(),
            _lib.ptr(loss_buf), B, N, d, m.num_class, red, LN_EPS, rdt, st), 'head_fwd')
        w.reduction = reduction
        loss = None
        if labels is not None:
            loss = w.loss_none if reduction == 'none' else w.loss[0]
        return loss, w.logits

    def _block_forward(self, l, w, record_attention=False, recompute=False):
        """block l: x[l] -> x[l + 1] (PreNorm attention + residual, PreNorm feed-forward + residual).
        recompute=True (activation checkpointing, from backward): rebuild what backward reads of the block -- ln1, qkv,
        lse, o, y, ln2, u, h -- from x[l]; the last Linear (whose output x[l + 1] is kept) is skipped.  Dropout masks are
        a function of (seed, site, element), so the recomputed tensors are bit-identical to the forward pass's."""
        m, lib, st = self.model, self.lib, self._stream
        c = m.config
        dt, rdt = m._dtype_code, m._res_code
        epi_res = EPI_BIAS_RES_F32 if m._res_f32 else EPI_BIAS_RES
        d, mlp, H = c.hidden_size, c.intermediate_size, c.num_attention_heads
        inner, dh = d, d // H
        B, N, M = w.B, w.N, w.M
        wt, pf = m._weights(), m._params_f32()
        p_blk = w.p_blk
        scale = float(dh) ** -0.5
        blk_seed = self.rng.data_ptr() if p_blk > 0 else None
        p = f'l{l}.'
        # dropout sites of block l: 1+4l attention probabilities, 2+4l after to_out, 3+4l after GELU, 4+4l after net[3]
        s_att, s_out, s_act, s_ff2 = 1 + 4 * l, 2 + 4 * l, 3 + 4 * l, 4 + 4 * l
        _lib.check(lib.ecgvit_layernorm_fwd(w.x[l].data_ptr(), pf[p + 'ln1.w'].data_ptr(), pf[p + 'ln1.b'].data_ptr(),
                                            w.ln1[l].data_ptr(), w.stat1[l][0].data_ptr(), w.stat1[l][1].data_ptr(),
                                            M, d, LN_EPS, rdt, st), 'layernorm_fwd')
        self._gemm(M, 3 * inner, d, w.ln1[l], d, 1, wt[p + 'qkv.w'], d, 1, EPI_STORE, w.qkv[l], 3 * inner)
        if record_attention:
            if l == 0:
                self.recorded_attention = torch.empty(B, c.num_hidden_layers, H, N, N, device=self.device,
                                                      dtype=torch.float32)
            rec = self.recorded_attention
            _lib.check(lib.ecgvit_attention_probs(w.qkv[l].data_ptr(), rec[:, l].data_ptr(), B, N, H, dh, scale,
                                                  rec.stride(0), dt, st), 'attention_probs')
        _lib.check(lib.ecgvit_attention_fwd(w.qkv[l].data_ptr(), w.o[l].data_ptr(), w.lse[l].data_ptr(), B, N, H, dh,
                                            scale, p_blk, s_att, blk_seed, dt, st), 'attention_fwd')
        self._gemm(M, d, inner, w.o[l], inner, 1, wt[p + 'out.w'], inner, 1, epi_res, w.y[l], d,
                   aux=w.x[l], bias=pf[p + 'out.b'], drop=(p_blk, s_out))
        _lib.check(lib.ecgvit_layernorm_fwd(w.y[l].data_ptr(), pf[p + 'ln2.w'].data_ptr(), pf[p + 'ln2.b'].data_ptr(),
                                            w.ln2[l].data_ptr(), w.stat2[l][0].data_ptr(), w.stat2[l][1].data_ptr(),
                                            M, d, LN_EPS, rdt, st), 'layernorm_fwd')
        self._gemm(M, mlp, d, w.ln2[l], d, 1, wt[p + 'ff1.w'], d, 1, EPI_BIAS_GELU, w.u[l], mlp,
                   out2=w.h[l], bias=pf[p + 'ff1.b'], drop=(p_blk, s_act))
        if not recompute:
            self._gemm(M, d, mlp, w.h[l], mlp, 1, wt[p + 'ff2.w'], mlp, 1, epi_res, w.x[l + 1], d,
                       aux=w.y[l], bias=pf[p + 'ff2.b'], drop=(p_blk, s_ff2))

    def span_buffer(self, B):
        """device int32 [B, 2] (start, length) of the TimeOut span of every record; a persistent buffer, so a captured
        step reads whatever `set_spans` wrote last"""
        buf = self._spans.get(B)
        if buf is None:
            buf = self._spans[B] = torch.zeros(B, 2, device=self.device, dtype=torch.int32)
        return buf

    def set_spans(self, spans):
        """spans: int [B, 2] (host or device), see `InputPipeline.draw_spans`"""
        self.span_buffer(spans.shape[0]).copy_(spans.to(torch.int32), non_blocking=True)

    def _loss_weight_table(self):
        """(device pointer, length) of `EcgVit.loss_weight` (ecg_vit.py:144-147), or (None, 0) when unset; the table
        is re-uploaded only when the attribute changes (never inside a captured step)"""
        lw = self.model.loss_weight
        if not lw:
            return None, 0
        key = tuple(float(v) for v in lw)
        if self._lw_key != key:
            self._lw_table = torch.tensor(key, dtype=torch.float32, device=self.device)
            self._lw_key = key
        return self._lw_table.data_ptr(), len(key)

    # ---- backward --------------------------------------------------------------------------------
    def backward(self, grad_scale=1.0, zero_grads=True, grad_scale_dev=None):
        """Gradients of the last forward's loss w.r.t. every parameter, accumulated into the flat grad buffer.
        grad_scale_dev: optional fp32 CUDA scalar multiplied into the upstream gradient on the device (no host sync)."""
        m, lib, st = self.model, self.lib, self._stream
        c = m.config
        dt = m._dtype_code
        rdt = m._res_code
        w = self._cur
        assert w is not None, 'backward before forward'
        assert w.reduction in ('mean', 'sum'), "backward needs loss_reduction 'mean' or 'sum'"
        B, n, N, M = w.B, w.n, w.N, w.M
        P, C, d, mlp, H = c.patch_size, c.num_channels, c.hidden_size, c.intermediate_size, c.num_attention_heads
        inner, dh = d, d // H
        depth = c.num_hidden_layers
        wt, pf, gr = m._weights(), m._params_f32(), m._grads_f32()
        if zero_grads:
            m._flat_g.zero_()
        p_emb, p_blk = w.p_emb, w.p_blk
        seed_ptr = self.rng.data_ptr()
        blk_seed = seed_ptr if p_blk > 0 else None
        last = f'l{depth - 1}.'
        top = (depth - 1) & 1
        # with dropout between net[3] / to_out[0] and the residual add, the Linear sees mask * dz / (1 - p): its operand
        # and its bias gradient come from the masked copy, so the producers must not pre-sum the unmasked gradient
        _lib.check(lib.ecgvit_head_bwd(
            w.x[depth].data_ptr(), pf['head.ln.w'].data_ptr(), pf['head.w'].data_ptr(), w.labels.data_ptr(),
            *self._loss_weight_table(), w.xn.data_ptr(), w.hstat[0].data_ptr(), w.hstat[1].data_ptr(), w.logits.data_ptr(),
            w.dz[top].data_ptr(),
            gr['head.w'].data_ptr(), gr['head.b'].data_ptr(), gr['head.ln.w'].data_ptr(), gr['head.ln.b'].data_ptr(),
            gr[last + 'ff2.b'].data_ptr() if p_blk == 0 else None, w.head_scratch.data_ptr(), B, N, d, m.num_class,
            _lib.REDUCTION[w.reduction], float(grad_scale), _lib.ptr(grad_scale_dev), rdt, st), 'head_bwd')
        scale = float(dh) ** -0.5

        main, side = torch.cuda.current_stream(), self.side_stream

        def wgrad(*args, **kw):
            """weight-gradient GEMM on the side stream, ordered after everything issued so far on the main stream;
            returns the event that marks its completion (None when there is no side stream)"""
            if side is None:
                self._gemm(*args, **kw)
                return None
            ready = torch.cuda.Event()
            ready.record(main)
            side.wait_event(ready)
            with torch.cuda.stream(side):
                self._gemm(*args, **kw)
                done = torch.cuda.Event()
                done.record(side)
            return done

        def before_overwrite(ev):
            """the main stream may not overwrite a scratch buffer that a side-stream GEMM is still reading"""
            if ev is not None:
                main.wait_event(ev)

        ev_fold = [None, None]

        def layernorm_bwd(which, dln, x_in, gamma, stat, dres, dx, dgamma, dbeta, dcol, dxm, p_drop, site):
            """dx = dres + LN'(dln) (+ the dropout-masked copy dxm) on the main stream; the fold of the per-CTA partial rows
            into dgamma / dbeta / dcol feeds nothing in the chain and runs on the side stream"""
            scr = w.ln_scratch[which]
            before_overwrite(ev_fold[which])  # the fold of this scratch's previous user
            seed = blk_seed if (dxm is not None and p_drop > 0) else None
            _lib.check(lib.ecgvit_layernorm_bwd(
                dln.data_ptr(), x_in.data_ptr(), gamma.data_ptr(), stat[0].data_ptr(), stat[1].data_ptr(),
                dres.data_ptr(), dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), _lib.ptr(dcol), scr.data_ptr(),
                _lib.ptr(dxm), p_drop, site, seed, M, d, 0 if side is None else 1, rdt, st),
                'layernorm_bwd' if side is None else 'layernorm_bwd_partial')
            if side is not None:
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    _lib.check(lib.ecgvit_layernorm_bwd_finalize(scr.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                                                 _lib.ptr(dcol), M, d, side.cuda_stream),
                               'layernorm_bwd_finalize')
                    ev_fold[which] = torch.cuda.Event()
                    ev_fold[which].record(side)

        # side-stream events per layer: ev[l] = {'ff2', 'ff1', 'out', 'qkv'} -> completion of that weight gradient
        evs = {}
        if p_blk > 0:
            # the top block's input gradient comes from the head, not from a LayerNorm': mask it here
            _lib.check(lib.ecgvit_dropout_bwd_copy(w.dz[top].data_ptr(), w.dzm[0][top].data_ptr(),
                                                   gr[last + 'ff2.b'].data_ptr(), M, d, d, p_blk, 4 * depth, seed_ptr, dt, st),
                       'dropout_bwd_copy')
        for l in range(depth - 1, -1, -1):
            p = f'l{l}.'
            q, qn = l & 1, (l - 1) & 1      # buffer parity of this layer / of the layer below
            s_att, s_out, s_act, s_ff2 = 1 + 4 * l, 2 + 4 * l, 3 + 4 * l, 4 + 4 * l
            old = evs.pop(l + 2, {})        # the previous users of this parity's buffers (a whole layer ago)
            if w.ckpt and l < depth - 2:
                # the activation set of layer l was last used by layer l + 2: its weight gradients must have read it
                for ev in old.values():
                    before_overwrite(ev)
                self._block_forward(l, w, recompute=True)
            e = evs.setdefault(l, {})
            dz, dy, du, dqkv = w.dz[q], w.dy[q], w.du[q], w.dqkv[q]
            dzl = w.dzm[0][q] if p_blk > 0 else dz      # what net[3] saw: mask * dz / (1 - p)
            # ---- feed-forward branch: x[l+1] = y + drop(W2 drop(gelu(W1 ln2(y) + b1)) + b2)
            e['ff2'] = wgrad(d, mlp, M, dzl, d, 0, w.h[l], mlp, 0, EPI_ATOMIC_F32, gr[p + 'ff2.w'], mlp, split_k=0)
            before_overwrite(old.get('ff1'))   # du of this parity
            self._gemm(M, mlp, d, dzl, d, 1, wt[p + 'ff2.w'], mlp, 0, EPI_DGELU, du, mlp, aux=w.u[l], drop=(p_blk, s_act))
            # FF1's bias gradient (column sums of du) feeds nothing in the chain: it rides on the side stream in front of
            # FF1's weight gradient (same operand, so the event that guards `du` covers both)
            if side is None:
                _lib.check(lib.ecgvit_colsum(du.data_ptr(), gr[p + 'ff1.b'].data_ptr(), M, mlp, mlp, dt, st), 'colsum')
            else:
                ready = torch.cuda.Event()
                ready.record(main)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    _lib.check(lib.ecgvit_colsum(du.data_ptr(), gr[p + 'ff1.b'].data_ptr(), M, mlp, mlp, dt,
                                                 side.cuda_stream), 'colsum')
            e['ff1'] = wgrad(mlp, d, M, du, mlp, 0, w.ln2[l], d, 0, EPI_ATOMIC_F32, gr[p + 'ff1.w'], d, split_k=0)
            self._gemm(M, d, mlp, du, mlp, 1, wt[p + 'ff1.w'], d, 0, EPI_STORE, w.dln, d)
            before_overwrite(old.get('out'))   # dy / its masked copy of this parity
            dyl = w.dzm[1][q] if p_blk > 0 else dy
            layernorm_bwd(0, w.dln, w.y[l], pf[p + 'ln2.w'], w.stat2[l], dz, dy, gr[p + 'ln2.w'], gr[p + 'ln2.b'],
                          gr[p + 'out.b'], dyl if p_blk > 0 else None, p_blk, s_out)
            # ---- attention branch: y = x + drop(Wo attn(Wqkv ln1(x)) + bo); dyl = mask * dy / (1 - p)
            e['out'] = wgrad(d, inner, M, dyl, d, 0, w.o[l], inner, 0, EPI_ATOMIC_F32, gr[p + 'out.w'], inner, split_k=0)
            self._gemm(M, inner, d, dyl, d, 1, wt[p + 'out.w'], inner, 0, EPI_STORE, w.d_o, inner)
            before_overwrite(old.get('qkv'))   # dqkv of this parity
            _lib.check(lib.ecgvit_attention_bwd(w.qkv[l].data_ptr(), w.o[l].data_ptr(), w.d_o.data_ptr(),
                                                w.lse[l].data_ptr(), dqkv.data_ptr(), _lib.ptr(w.attn_scratch), B, N,
                                                H, dh, scale, p_blk, s_att, blk_seed, dt, st),
                       'attention_bwd' if w.attn_scratch is None else 'attention_bwd_flash')
            e['qkv'] = wgrad(3 * inner, d, M, dqkv, 3 * inner, 0, w.ln1[l], d, 0, EPI_ATOMIC_F32, gr[p + 'qkv.w'], d,
                             split_k=0)
            self._gemm(M, d, 3 * inner, dqkv, 3 * inner, 1, wt[p + 'qkv.w'], d, 0, EPI_STORE, w.dln, d)
            below_bias = gr[f'l{l - 1}.ff2.b'] if l > 0 else None
            drop_below = p_blk > 0 and l > 0
            # LayerNorm' writes the input gradient of layer l - 1 into the OTHER parity, last read by the ff2 weight
            # gradient of layer l + 1
            before_overwrite(evs.get(l + 1, {}).get('ff2'))
            layernorm_bwd(1, w.dln, w.x[l], pf[p + 'ln1.w'], w.stat1[l], dy, w.dz[qn], gr[p + 'ln1.w'], gr[p + 'ln1.b'],
                          below_bias, w.dzm[0][qn] if drop_below else None, p_blk if drop_below else 0.0, s_ff2 - 4)
            if m._after_layer_backward is not None:
                # the gradient bucket of layer l is complete once its side-stream wgrads AND everything the main stream
                # has issued so far (LayerNorm / bias gradients) are done.  The collective is therefore issued from the
                # side stream after it has waited for the main stream's current point: c10d orders the NCCL kernel
                # after the issuing stream, and the main chain never waits for the weight-gradient GEMMs.
                if side is None:
                    m._after_layer_backward(l)
                else:
                    here = torch.cuda.Event()
                    here.record(main)
                    side.wait_event(here)
                    with torch.cuda.stream(side):
                        m._after_layer_backward(l)
        for e in evs.values():
            for ev in e.values():
                before_overwrite(ev)  # join the side stream
        for ev in ev_fold:
            before_overwrite(ev)
        dz = w.dz[(-1) & 1]   # gradient w.r.t. the token matrix (input of block 0)
        # ---- embedding: tok = [cls | a_patch We^T + be] + pos
        _lib.check(lib.ecgvit_embed_assemble_bwd(dz.data_ptr(), w.de.data_ptr(), gr['cls'].data_ptr(),
                                                 gr['pos'].data_ptr(), gr['embed.b'].data_ptr(), B, n, d, p_emb, 0,
                                                 seed_ptr if p_emb > 0 else None, dt, st), 'embed_assemble_bwd')
        if w.Kp == w.K:
            self._gemm(d, w.K, B * n, w.de, d, 0, w.a_patch, w.K, 0, EPI_ATOMIC_F32, gr['embed.w'], w.K, split_k=0)
        else:
            w.embed_gpad.zero_()
            self._gemm(d, w.Kp, B * n, w.de, d, 0, w.a_patch, w.Kp, 0, EPI_ATOMIC_F32, w.embed_gpad, w.Kp, split_k=0)
            _lib.check(lib.ecgvit_unpad_add_f32(w.embed_gpad.data_ptr(), gr['embed.w'].data_ptr(), d, w.K, w.Kp, st),
                       'unpad_add')
        if m._after_layer_backward is not None:
            m._after_layer_backward(-1)
