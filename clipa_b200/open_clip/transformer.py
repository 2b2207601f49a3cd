"""Transformer towers with the reference's module/parameter names (open_clip/transformer.py) and
the B200 kernels underneath.

The nn.Module tree (and therefore the state_dict keys, the weight-decay split on parameter names,
DDP wrapping, checkpoints) is the reference's: ln_1 / attn.in_proj_weight / attn.out_proj /
ln_2 / mlp.c_fc / mlp.c_proj per block.  torch.nn modules are used as PARAMETER CONTAINERS with
their stock initialisation; their forward() is never called -- each block is one
clipa_b200.functional.ResidualBlockFn node.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional, Tuple

import torch
from torch import nn

from .. import functional as Fn
from .. import ops
from .._lib import POOL_ARGMAX_ID, POOL_FIRST, POOL_LAST, POOL_MEAN_ALL, POOL_MEAN_SKIP_FIRST
from ..ops import ACT_BY_NAME
from .pos_embed import get_2d_sincos_pos_embed


SUPPORTED_HEAD_DIMS = (64, 80, 96, 128)      # csrc/attention*.cu


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class LayerNorm(nn.LayerNorm):
    """open_clip/transformer.py:19-34 (LayerNorm and LayerNormFp32 collapse into one kernel:
    statistics are always fp32, output is cast back to the input dtype).  x: [..., D] bf16."""

    def forward(self, x: torch.Tensor):
        shape = x.shape
        y = Fn.LayerNormFn.apply(x.reshape(-1, shape[-1]).contiguous(), self.weight, self.bias, self.eps)
        return y.reshape(shape)


LayerNormFp32 = LayerNorm


class PatchDropout(nn.Module):
    """Random token dropping of the image sequence in training (open_clip/transformer.py:53-90,
    arXiv 2212.00794): keeps `max(1, int(num_tokens * (1 - prob)))` patch tokens per sample, chosen as
    the top-k of i.i.d. normal scores, in top-k ORDER (the reference does not re-sort them; positional
    embeddings are already added, and the non-causal tower is permutation-equivariant), with the CLS
    token kept in front.  Identity in eval mode or for prob == 0.

    Unlike the reference, which draws the scores with a default-device `torch.randn` (a host round trip
    when the activations live on the GPU, transformer.py:76,82), scores are drawn on the activations'
    device.  `score_fn(batch, num_tokens, device)` can be replaced to inject scores (parity tests)."""

    def __init__(self, prob: float, exclude_first_token: bool = True):
        super().__init__()
        assert 0 <= prob < 1.
        self.prob = prob
        self.exclude_first_token = exclude_first_token
        self.score_fn = lambda batch, num_tokens, device: torch.randn(batch, num_tokens, device=device)

    def forward(self, x: torch.Tensor) -> torch.Tensor:      # [batch, tokens, width]
        if not self.training or self.prob == 0.:
            return x
        if self.exclude_first_token:
            cls_tokens, x = x[:, :1], x[:, 1:]
        batch, num_tokens = x.shape[0], x.shape[1]
        keep = max(1, int(num_tokens * (1 - self.prob)))
        idx = self.score_fn(batch, num_tokens, x.device).topk(keep, dim=-1).indices
        x = torch.gather(x, 1, idx.unsqueeze(-1).expand(batch, keep, x.shape[2]))
        if self.exclude_first_token:
            x = torch.cat((cls_tokens, x), dim=1)
        return x


class ResidualAttentionBlock(nn.Module):
    """open_clip/transformer.py:195-250 (self-attention form; ls_1/ls_2 are Identity as in every
    shipped CLIPA config)."""

    def __init__(self, d_model: int, n_head: int, mlp_ratio: float = 4.0, ls_init_value: float = None,
                 act: str = "gelu", is_cross_attention: bool = False):
        super().__init__()
        if ls_init_value is not None or is_cross_attention:
            raise NotImplementedError("LayerScale / cross-attention blocks are outside the CLIPA hot path")
        if d_model % n_head or d_model // n_head not in SUPPORTED_HEAD_DIMS:
            raise NotImplementedError(
                f"head_dim {d_model / n_head:g} (width {d_model}, {n_head} heads): attention kernels exist for head_dim 64 "
                "and 80 (tcgen05, any sequence length) and 96 / 128 (mma.sync, sequences up to ~380 tokens); "
                "ViT-g/14 (88) and ViT-bigG/14 (104) are not built")
        self.ln_1 = LayerNorm(d_model)
        self.attn = nn.MultiheadAttention(d_model, n_head)   # parameter container only
        self.ln_2 = LayerNorm(d_model)
        mlp_width = int(d_model * mlp_ratio)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, mlp_width)),
            ("gelu", nn.Identity()),                           # activation is fused into the c_fc GEMM
            ("c_proj", nn.Linear(mlp_width, d_model)),
        ]))
        self.n_head = n_head
        self.act = ACT_BY_NAME[act]

    def forward(self, x: torch.Tensor, batch: int, seq: int, causal: bool, save_ln: bool = False,
                prev_proj_bias: Optional[torch.Tensor] = None, proj_bias_grad_by_next: bool = False,
                recompute_attn_out: bool = False, keep_mlp: bool = False):
        """x: [batch*seq, d_model] bf16, sample-major rows.  prev_proj_bias / proj_bias_grad_by_next: the c_proj bias
        gradient of a block is the column sum of the gradient the NEXT block's LayerNorm backward writes, so the next
        block produces it (see functional.ResidualBlockFn)."""
        return Fn.ResidualBlockFn.apply(
            x, self.ln_1.weight, self.ln_1.bias, self.attn.in_proj_weight, self.attn.in_proj_bias,
            self.attn.out_proj.weight, self.attn.out_proj.bias, self.ln_2.weight, self.ln_2.bias,
            self.mlp.c_fc.weight, self.mlp.c_fc.bias, self.mlp.c_proj.weight, self.mlp.c_proj.bias,
            batch, seq, self.n_head, causal, self.act, save_ln, prev_proj_bias, proj_bias_grad_by_next,
            recompute_attn_out, keep_mlp)


class _GradReady:
    """Tensor hook: reports, does not touch the gradient."""

    def __init__(self, cb, tower, index):
        self.cb, self.tower, self.index = cb, tower, index

    def __call__(self, grad):
        self.cb(self.tower, self.index)
        return None


def activation_levels(free: int, unit: int, layers: int, mlp_units: float, keep_reserve: int = 0,
                      save_ln="auto", drop_o="auto", keep_mlp="auto"):
    """Activation-memory level of a tower from the HBM that is free for it (bytes, already minus what the rest of the
    step still needs), `unit` = bytes of one [tokens, width] bf16 tensor.  -> (save LayerNorm outputs, recompute the
    attention output, number of trailing blocks that keep their MLP activations).  "auto" entries are decided here,
    anything else is forced (see the comment block in Transformer)."""
    if save_ln == "auto":
        save_ln = (8 * layers + 10) * unit + (6 << 30) < free
    else:
        save_ln = bool(save_ln)
    if drop_o == "auto":
        drop_o = (not save_ln) and (6 * layers + 10) * unit + (3 << 30) > free
    else:
        drop_o = bool(drop_o)
    if keep_mlp == "auto":
        spare = free - (8 * layers + 10) * unit - (8 << 30) - keep_reserve
        keep = int(spare // (mlp_units * unit)) if (save_ln and spare > 0) else 0
    else:
        keep = int(keep_mlp)
    return save_ln, drop_o, max(0, min(layers, keep))


class Transformer(nn.Module):
    """open_clip/transformer.py:294-326.  `grad_checkpointing` is accepted for API parity; the block
    function already recomputes its MLP activations, so the flag does not change the math."""

    def __init__(self, width: int, layers: int, heads: int, mlp_ratio: float = 4.0,
                 ls_init_value: float = None, act: str = "gelu"):
        super().__init__()
        self.width = width
        self.layers = layers
        self.grad_checkpointing = False
        self.resblocks = nn.ModuleList([
            ResidualAttentionBlock(width, heads, mlp_ratio, ls_init_value=ls_init_value, act=act)
            for _ in range(layers)])

    def get_cast_dtype(self) -> torch.dtype:
        return self.resblocks[0].mlp.c_fc.weight.dtype

    # Activation memory policy per tower, chosen from the free HBM at forward time ('auto'), one "unit" = one
    # [batch*L, width] bf16 tensor (backward adds ~10 transient units: f, g / df, dh, dx):
    #   8 units / block  x, qkv (3), o, x1 + the two LayerNorm outputs     when it fits with a 6 GiB margin
    #   6 units / block  LayerNorm outputs recomputed in backward          default
    #   5 units / block  attention output o recomputed too (one extra attention forward per block)
    #                    when even 6 units + 3 GiB would not fit (ViT-H/14 at 8192 pairs per GPU: 2 GiB were left)
    #   + 2 x mlp_ratio units for every block that also keeps act(f) and act'(f) of its MLP (no c_fc recompute GEMM in
    #                    backward: -10 % of a block's GEMM FLOPs); as many of the LAST blocks as fit next to the
    #                    8-unit level with an 8 GiB margin -- all of them at the 2048-pair micro-batches of a
    #                    GradCache schedule, a handful at 4096 pairs per GPU
    # `other_need_bytes` = what still has to fit AFTER this tower's forward (the other tower, the contrastive head):
    # set by CLIP.forward.  save_ln_outputs / recompute_attn_out = True / False force a level; keep_mlp_blocks =
    # an int forces the number of MLP-keeping blocks.
    save_ln_outputs = "auto"
    recompute_attn_out = "auto"
    keep_mlp_blocks = "auto"
    other_need_bytes = 0
    keep_reserve_bytes = 0     # extra room the MLP-keeping decision leaves for the tower that runs afterwards

    def base_need_bytes(self, rows: int) -> int:
        """Bytes this tower keeps for backward at the default level (+ backward transients), for `rows` tokens."""
        return (6 * self.layers + 10) * rows * self.width * 2

    def full_need_bytes(self, rows: int) -> int:
        """Same at the richest level: LayerNorm outputs and the MLP activations of every block kept."""
        mlp_units = 2 * self.resblocks[0].mlp.c_fc.out_features / self.width
        return int(((8 + mlp_units) * self.layers + 10) * rows * self.width * 2)

    def _activation_policy(self, x: torch.Tensor):
        """-> (save_ln, drop_o, number of trailing blocks that keep their MLP activations)"""
        if not torch.is_grad_enabled() or not x.is_cuda:
            return False, False, 0
        free, _ = torch.cuda.mem_get_info(x.device)
        free += torch.cuda.memory_reserved(x.device) - torch.cuda.memory_allocated(x.device)
        return activation_levels(free - self.other_need_bytes, x.numel() * x.element_size(), self.layers,
                                 2 * self.resblocks[0].mlp.c_fc.out_features / self.width, self.keep_reserve_bytes,
                                 self.save_ln_outputs, self.recompute_attn_out, self.keep_mlp_blocks)

    def _decide_save_ln(self, x: torch.Tensor) -> bool:
        return self._activation_policy(x)[0]

    # TrainStep's overlapped gradient all-reduce: callback(tower, i) fires in backward once the gradients of every
    # block >= i are complete (tensor hook on the input of block i), for i = 0, k, 2k, ...
    grad_ready_callback = None
    grad_bucket_blocks = 4

    def forward(self, x: torch.Tensor, batch: int, seq: int, causal: bool = False):
        save_ln, drop_o, keep = self._activation_policy(x)
        if torch.is_grad_enabled():
            self.last_policy = (save_ln, drop_o, keep)      # reported by bench.py
        cb = self.grad_ready_callback if torch.is_grad_enabled() else None
        n_blocks = len(self.resblocks)
        for i, r in enumerate(self.resblocks):
            if cb is not None and x.requires_grad and i % self.grad_bucket_blocks == 0:
                x.register_hook(_GradReady(cb, self, i))
            prev = self.resblocks[i - 1].mlp.c_proj.bias if i > 0 else None
            x = r(x, batch, seq, causal, save_ln, prev, i + 1 < n_blocks, drop_o, i >= n_blocks - keep)
        return x


class VisionTransformer(nn.Module):
    """open_clip/transformer.py:329-534 (no attentional pool, no input patchnorm)."""

    def __init__(self, image_size: int, patch_size: int, width: int, layers: int, heads: int,
                 mlp_ratio: float, ls_init_value: float = None, global_average_pool: bool = False,
                 output_dim: int = 512, patch_dropout: float = 0., act: str = "gelu",
                 pos_embed: str = "learnable", ln_pre: bool = True, pool_style: str = "open_clip",
                 output_tokens: bool = False):
        super().__init__()
        # a patch_dropout of 0. means disabled (open_clip/transformer.py:387-388)
        self.patch_dropout = PatchDropout(patch_dropout) if patch_dropout > 0. else nn.Identity()
        self.output_tokens = output_tokens
        image_height, image_width = self.image_size = to_2tuple(image_size)
        patch_height, patch_width = self.patch_size = to_2tuple(patch_size)
        self.grid_size = (image_height // patch_height, image_width // patch_width)
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)  # container
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        n_pos = self.grid_size[0] * self.grid_size[1] + 1
        if pos_embed == "learnable":
            self.positional_embedding = nn.Parameter(scale * torch.randn(n_pos, width))
        elif pos_embed == "sin_cos_2d":
            assert self.grid_size[0] == self.grid_size[1], "sin-cos 2d position embedding needs a square grid"
            self.positional_embedding = nn.Parameter(
                get_2d_sincos_pos_embed(width, self.grid_size[0], cls_token=True), requires_grad=False)
        else:
            raise NotImplementedError(pos_embed)
        self.ln_pre = LayerNorm(width) if ln_pre else nn.Identity()
        self.transformer = Transformer(width, layers, heads, mlp_ratio, ls_init_value=ls_init_value, act=act)
        self.global_average_pool = global_average_pool
        self.attn_pool = None
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self.pool_style = pool_style
        self.width = width

    def lock(self, unlocked_groups=0, freeze_bn_stats=False):
        """LiT-style freezing, open_clip/transformer.py:415-446."""
        for p in self.parameters():
            p.requires_grad = False
        if unlocked_groups != 0:
            groups = [[self.conv1, self.class_embedding, self.positional_embedding, self.ln_pre],
                      *self.transformer.resblocks[:-1],
                      [self.transformer.resblocks[-1], self.ln_post], self.proj]

            def _unlock(x):
                if isinstance(x, (list, tuple)):
                    for g in x:
                        _unlock(g)
                elif isinstance(x, nn.Parameter):
                    x.requires_grad = True
                else:
                    for p in x.parameters():
                        p.requires_grad = True
            _unlock(groups[-unlocked_groups:])

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.transformer.grad_checkpointing = enable

    def _global_pool(self, x: torch.Tensor, include_cls=True) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.global_average_pool and not include_cls:
            return x[:, 1:].float().mean(dim=1).to(x.dtype), x[:, 1:]
        if self.global_average_pool and include_cls:
            return x.float().mean(dim=1).to(x.dtype), x
        return x[:, 0], x[:, 1:]

    def forward(self, x: torch.Tensor):
        N = x.shape[0]
        ph, pw = self.patch_size
        gh, gw = self.grid_size
        W = self.width
        # conv1 (stride == kernel, no bias) == patchify + GEMM (open_clip/transformer.py:371,491-493).
        # K = 3*ph*pw is zero-padded to a multiple of 8 so TMA row pitches stay 16-byte aligned.
        K = 3 * ph * pw
        Kp = (K + 7) // 8 * 8
        if x.dtype not in (torch.bfloat16, torch.float32):
            x = x.float()
        patches = ops.patchify(x.contiguous(), ph, pw, Kp)          # [N*gh*gw, Kp] bf16, one pass
        wmat = self.conv1.weight.reshape(W, K)
        if Kp != K:
            wmat = torch.nn.functional.pad(wmat, (0, Kp - K))
        tok = Fn.LinearFn.apply(patches, wmat, None, False)
        L = gh * gw + 1
        dropping = isinstance(self.patch_dropout, PatchDropout) and self.training and self.patch_dropout.prob > 0.
        x = Fn.AssembleTokensFn.apply(tok, self.class_embedding, self.positional_embedding, N, L)   # [N, L, W]
        if dropping:
            x = self.patch_dropout(x)      # after the positional embedding, before ln_pre (transformer.py:501-502)
            L = x.shape[1]
        x = self.ln_pre(x)
        x = self.transformer(x.reshape(N * L, W).contiguous(), N, L, causal=False).reshape(N, L, W)
        if self.output_tokens:
            return self._forward_tail_with_tokens(x)
        # pooling is a row selection / token mean, LayerNorm is row-wise: ln_post runs on the pooled rows only
        # (big_vision_tok applies it to every token first, transformer.py:518-520 -- same pooled rows)
        if self.pool_style in ("open_clip", "big_vision_tok"):
            if self.pool_style == "big_vision_tok":
                assert not self.global_average_pool
            mode = POOL_MEAN_ALL if self.global_average_pool else POOL_FIRST
        elif self.pool_style == "big_vision_gap":
            assert self.global_average_pool
            mode = POOL_MEAN_SKIP_FIRST
        else:
            raise ValueError(self.pool_style)
        pooled = self.ln_post(Fn.PoolTokensFn.apply(x, mode, None))
        if self.proj is not None:
            pooled = Fn.LinearFn.apply(pooled.contiguous(), self.proj, None, True)
        return pooled

    def _forward_tail_with_tokens(self, x: torch.Tensor):
        """output_tokens=True (CoCa-style consumers): pooled features AND the token sequence, torch indexing."""
        if self.pool_style == "open_clip":
            pooled, tokens = self._global_pool(x)
            pooled = self.ln_post(pooled)
        elif self.pool_style == "big_vision_tok":
            assert not self.global_average_pool
            x = self.ln_post(x)
            pooled, tokens = self._global_pool(x)
        elif self.pool_style == "big_vision_gap":
            assert self.global_average_pool
            pooled, tokens = self._global_pool(x, include_cls=False)
            pooled = self.ln_post(pooled)
        else:
            raise ValueError(self.pool_style)
        if self.proj is not None:
            pooled = Fn.LinearFn.apply(pooled.contiguous(), self.proj, None, True)
        return pooled, tokens


class TextTransformer(nn.Module):
    """Parameter layout and init of open_clip/transformer.py:537-616; CLIP.__init__ re-homes these
    members onto the CLIP module exactly like the reference (open_clip/model.py:216-225)."""

    def __init__(self, context_length: int = 77, vocab_size: int = 49408, width: int = 512,
                 heads: int = 8, layers: int = 12, ls_init_value: float = None, output_dim: int = 512,
                 act: str = "gelu", embed_cls: bool = False, pad_id: int = 0,
                 pool_style: str = "open_clip", attention_mask: bool = True):
        super().__init__()
        if embed_cls:
            raise NotImplementedError("embed_cls (CoCa text tower) is outside the CLIPA hot path")
        self.num_pos = self.context_length = context_length
        self.vocab_size = vocab_size
        self.width = width
        self.output_dim = output_dim
        self.heads = heads
        self.pad_id = pad_id
        self.pool_style = pool_style
        self.text_projection = nn.Parameter(torch.empty(width, output_dim))
        self.token_embedding = nn.Embedding(vocab_size, width)
        self.positional_embedding = nn.Parameter(torch.empty(self.num_pos, width))
        self.transformer = Transformer(width=width, layers=layers, heads=heads,
                                       ls_init_value=ls_init_value, act=act)
        self.ln_final = LayerNorm(width)
        if attention_mask:
            mask = torch.full((self.num_pos, self.num_pos), float("-inf")).triu_(1)
            self.register_buffer("attn_mask", mask, persistent=False)
        else:
            self.attn_mask = None
        self.init_parameters()

    def init_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def lock(self, unlocked_layers: int = 0, freeze_layer_norm: bool = True):
        """Freeze the text tower, optionally leaving the last `unlocked_layers` blocks (and the final LayerNorm /
        projection) trainable -- the role of TextTransformer.lock for CustomTextCLIP.lock_text_tower
        (open_clip/model.py:304-305)."""
        for p in self.parameters():
            p.requires_grad = False
        if unlocked_layers > 0:
            for blk in self.transformer.resblocks[-unlocked_layers:]:
                for p in blk.parameters():
                    p.requires_grad = True
            self.text_projection.requires_grad = True
            for p in self.ln_final.parameters():
                p.requires_grad = not freeze_layer_norm

    @torch.jit.ignore
    def set_grad_checkpointing(self, enable=True):
        self.transformer.grad_checkpointing = enable

    def forward(self, text: torch.Tensor):
        """open_clip/transformer.py:638-681 (no cls_emb): token gather + positional rows [:seq_len] -> blocks ->
        ln_final -> pool (EOT argmax / first / last) -> text_projection.  Used by CustomTextCLIP."""
        return text_tower_forward(self, text, slice_positions=True)


def text_tower_forward(mod, text: torch.Tensor, slice_positions: bool):
    """The text tower on the sm_100a kernels, shared by CLIP.encode_text (whose members live on the CLIP
    module itself, open_clip/model.py:216-225,245-263) and TextTransformer.forward (transformer.py:638-681):
    `mod` provides token_embedding, positional_embedding, transformer, ln_final, text_projection, attn_mask,
    pool_style, context_length.  CLIP.encode_text adds the WHOLE positional table (model.py:247: the text
    length must equal context_length); TextTransformer.forward slices it to the sequence length."""
    N, L = text.shape
    if not slice_positions and L != mod.context_length:
        raise ValueError(f"text length {L} != context_length {mod.context_length} "
                         "(the reference adds the full positional embedding, model.py:247)")
    if L > mod.positional_embedding.shape[0]:
        raise ValueError(f"text length {L} exceeds the {mod.positional_embedding.shape[0]} positions of the tower")
    x = Fn.EmbedTokensFn.apply(text, mod.token_embedding.weight, mod.positional_embedding)   # gather + positions
    W = x.shape[-1]
    causal = mod.attn_mask is not None
    x = mod.transformer(x.reshape(N * L, W), N, L, causal=causal).reshape(N, L, W)
    # ln_final is row-wise, so normalising only the pooled rows equals model.py:251-254
    if mod.pool_style == 'open_clip':
        pooled = Fn.PoolTokensFn.apply(x, POOL_ARGMAX_ID, text)       # EOT = first position of the largest id
    elif mod.pool_style == 'big_vision_tok':
        pooled = Fn.PoolTokensFn.apply(x, POOL_FIRST, None)
    elif mod.pool_style == 'big_vision_last':
        pooled = Fn.PoolTokensFn.apply(x, POOL_LAST, None)
    else:
        raise ValueError(mod.pool_style)
    pooled = mod.ln_final(pooled)
    if mod.text_projection is not None:
        pooled = Fn.LinearFn.apply(pooled, mod.text_projection, None, True)
    return pooled
