"""
ORACLE (test infrastructure, NOT product code) -- CPU restatement of `vit_pytorch.ViT`.

The reference (`/root/reference/ecg_transformer/models/ecg_vit.py:12-13,116,141`) delegates all
arithmetic of the ECG-ViT forward pass to the un-vendored PyPI package `vit-pytorch==0.33.2`
(`/root/reference/requirements.txt:174`).  That package is absent from /root/reference and cannot be
installed offline, so its published algorithm is restated here in plain PyTorch so that

  * the reference's own `EcgVit` wrapper can be imported verbatim on top of it (`oracle/ref_shim.py`),
  * the module tree reproduces the `state_dict` keys a reference checkpoint carries
    (`ecg_vit.py:159` loads with `strict=True`):
        pos_embedding, cls_token, to_patch_embedding.1.{weight,bias},
        transformer.layers.{i}.0.norm.*, transformer.layers.{i}.0.fn.to_qkv.weight,
        transformer.layers.{i}.0.fn.to_out.0.*, transformer.layers.{i}.1.norm.*,
        transformer.layers.{i}.1.fn.net.{0,3}.*, mlp_head.{0,1}.*

PARITY UNPINNED: the reference ships no tests / golden vectors / checkpoints for this path
(SURVEY.md section 4, 8c), so this restatement is anchored on the reference's call sites only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import torch
from torch import nn
from einops import rearrange, repeat
from einops.layers.torch import Rearrange


def _as_pair(v):
    return v if isinstance(v, tuple) else (v, v)


class PreNorm(nn.Module):
    """LayerNorm (eps 1e-5, affine) in front of a wrapped sub-layer; keys `norm.*`, `fn.*`."""

    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kwargs):
        return self.fn(self.norm(x), **kwargs)


class FeedForward(nn.Module):
    """Linear -> exact (erf) GELU -> Dropout -> Linear -> Dropout; keys `net.0.*`, `net.3.*`."""

    def __init__(self, dim, hidden_dim, dropout=0.):
        super().__init__()
        self.net = nn.Sequential(
            nn.Linear(dim, hidden_dim),
            nn.GELU(),
            nn.Dropout(dropout),
            nn.Linear(hidden_dim, dim),
            nn.Dropout(dropout),
        )

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    """Multi-head softmax attention with one bias-free packed q|k|v projection.

    `attend` (the softmax) is a sub-module because `vit_pytorch.recorder.Recorder` hooks it
    (reference use: `ecg_vit.py:176-180`).
    """

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.):
        super().__init__()
        inner = dim_head * heads
        needs_projection = not (heads == 1 and dim_head == dim)
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.attend = nn.Softmax(dim=-1)
        self.dropout = nn.Dropout(dropout)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout)) if needs_projection \
            else nn.Identity()

    def forward(self, x):
        q, k, v = (rearrange(t, 'b n (h d) -> b h n d', h=self.heads) for t in self.to_qkv(x).chunk(3, dim=-1))
        scores = torch.matmul(q, k.transpose(-1, -2)) * self.scale
        probs = self.dropout(self.attend(scores))
        out = rearrange(torch.matmul(probs, v), 'b h n d -> b n (h d)')
        return self.to_out(out)


class Transformer(nn.Module):
    """depth x [PreNorm(Attention) + residual, PreNorm(FeedForward) + residual]; no final norm."""

    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout=0.):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([
                PreNorm(dim, Attention(dim, heads=heads, dim_head=dim_head, dropout=dropout)),
                PreNorm(dim, FeedForward(dim, mlp_dim, dropout=dropout)),
            ]) for _ in range(depth)
        ])

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return x


class ViT(nn.Module):
    """Patchify (time-major, channel-minor features) -> Linear -> [CLS | patches] + pos -> blocks -> head."""

    def __init__(self, *, image_size, patch_size, num_classes, dim, depth, heads, mlp_dim, pool='cls',
                 channels=3, dim_head=64, dropout=0., emb_dropout=0.):
        super().__init__()
        img_h, img_w = _as_pair(image_size)
        p_h, p_w = _as_pair(patch_size)
        assert img_h % p_h == 0 and img_w % p_w == 0, 'Image dimensions must be divisible by the patch size.'
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        n_patch = (img_h // p_h) * (img_w // p_w)
        patch_dim = channels * p_h * p_w

        self.to_patch_embedding = nn.Sequential(
            Rearrange('b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p1=p_h, p2=p_w),
            nn.Linear(patch_dim, dim),
        )
        self.pos_embedding = nn.Parameter(torch.randn(1, n_patch + 1, dim))
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, num_classes))

    def forward(self, img):
        x = self.to_patch_embedding(img)
        b, n, _ = x.shape
        x = torch.cat((repeat(self.cls_token, '() n d -> b n d', b=b), x), dim=1)
        x += self.pos_embedding[:, :(n + 1)]
        x = self.dropout(x)
        x = self.transformer(x)
        x = x.mean(dim=1) if self.pool == 'mean' else x[:, 0]
        return self.mlp_head(self.to_latent(x))


class Recorder(nn.Module):
    """Restated `vit_pytorch.recorder.Recorder`: hooks every Attention.attend output.

    forward(img) -> (logits, attns[b, layers, heads, n, n]).  Reference use: `ecg_vit.py:176-193`.
    """

    def __init__(self, vit, device=None):
        super().__init__()
        self.vit = vit
        self.data = None
        self.recordings = []
        self.hooks = []
        self.hook_registered = False
        self.ejected = False
        self.device = device

    def _hook(self, _, inp, output):
        self.recordings.append(output.clone().detach())

    def _register_hook(self):
        for m in self.vit.transformer.modules():
            if isinstance(m, Attention):
                self.hooks.append(m.attend.register_forward_hook(self._hook))
        self.hook_registered = True

    def eject(self):
        self.ejected = True
        for h in self.hooks:
            h.remove()
        self.hooks.clear()
        return self.vit

    def clear(self):
        self.recordings.clear()

    def forward(self, img):
        assert not self.ejected, 'recorder has been ejected, cannot be used anymore'
        self.clear()
        if not self.hook_registered:
            self._register_hook()
        pred = self.vit(img)
        target = self.device if self.device is not None else img.device
        recs = tuple(r.to(target) for r in self.recordings)
        attns = torch.stack(recs, dim=1) if len(recs) > 0 else None
        return pred, attns
