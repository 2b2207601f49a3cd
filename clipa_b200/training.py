"""One optimizer step of CLIPA training, mirroring `train_one_epoch`'s inner loop
(clipa_torch/training/train.py:180-286) on top of clipa_b200.open_clip:

  H2D copy -> uint8 -> float/255 -> Normalize -> bf16 (`--to-float-on-device`, train.py:191-197)
  -> forward -> ClipLoss (local_loss + gather_with_grad) -> backward -> gradient all-reduce (mean)
  -> AdamW (two weight-decay groups, main.py:311-326) -> logit_scale clamp (train.py:285-286).

When the per-rank batch exceeds `micro_batch`, the GradCache schedule of the reference's
accum_freq > 1 path (train.py:216-256) is used: features of all chunks are first computed without
autograd, the loss and d(features) are evaluated once on the whole (gathered) batch, then every
chunk is re-run with autograd and back-propagated from its slice of d(features) -- except the last
chunk of the feature pass, whose graph is kept, so N chunks cost 2N-1 forwards.

Data parallelism: one process per GPU.  Parameters' gradients live in flat fp32 buffers (one per
weight-decay group), so the data-parallel reduction is two NCCL all-reduces (1.7 GB for ViT-L/14,
~4 ms on NVLink 5) issued after backward; at ~0.8 s per step it needs no bucketing or overlap.  The
optimizer is the fused clipa_adamw_step kernel over the same flat buffers (update + bf16 shadow
weights + gradient clear + 1/world_size averaging in one pass).
"""
from __future__ import annotations

import contextlib
import math
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import open_clip


def exclude_from_wd(name: str, p: torch.Tensor) -> bool:
    """training/main.py:311: no weight decay on gains, biases, LayerNorm, logit_scale."""
    return p.ndim < 2 or "bn" in name or "ln" in name or "bias" in name or "logit_scale" in name


class TrainStep:
    _armed = False          # overlapped all-reduce in flight for the current backward (see _setup_overlap)

    def __init__(self, model: torch.nn.Module, *, rank: int = 0, world_size: int = 1,
                 lr: float = 1.024e-3, betas: Tuple[float, float] = (0.9, 0.95), eps: float = 1e-6,
                 wd: float = 0.2, micro_batch: int = 4096, local_loss: bool = True,
                 gather_with_grad: bool = True, image_mean=None, image_std=None,
                 grad_clip_norm: Optional[float] = None, fused_optimizer: bool = True,
                 reference_accum_logit_scale: bool = True, overlap_grad_allreduce: bool = False,
                 allreduce_bucket_blocks: int = 4):
        self.model = model
        self.rank, self.world_size = rank, world_size
        self.micro_batch = micro_batch
        self.grad_clip_norm = grad_clip_norm
        # train.py:243-256 calls backward() once per accumulation chunk on the FULL loss, so the reference's
        # logit_scale gradient is accumulated accum_freq times (tower parameters are not: chunk j only reaches
        # them through its own features).  Reproduced by default; False gives the single-pass gradient.
        self.reference_accum_logit_scale = reference_accum_logit_scale
        self.overlap = overlap_grad_allreduce and world_size > 1
        self._bucket_blocks = allreduce_bucket_blocks
        self._armed = False
        self._works, self._done_ranges = [], []
        self.loss_fn = open_clip.ClipLoss(local_loss=local_loss, gather_with_grad=gather_with_grad,
                                          cache_labels=True, rank=rank, world_size=world_size)
        dev = next(model.parameters()).device
        self.device = dev
        mean = image_mean or getattr(model.visual, "image_mean", open_clip.factory.OPENAI_DATASET_MEAN)
        std = image_std or getattr(model.visual, "image_std", open_clip.factory.OPENAI_DATASET_STD)
        self._mean_host, self._std_host = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        self._mean = torch.tensor(mean, device=dev, dtype=torch.float32).reshape(1, 3, 1, 1)
        self._inv_std = (1.0 / torch.tensor(std, device=dev, dtype=torch.float32)).reshape(1, 3, 1, 1)
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.params = [p for _, p in named]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.step_count = 0
        gain = [p for n, p in named if exclude_from_wd(n, p)]
        rest = [p for n, p in named if not exclude_from_wd(n, p)]
        self.fused = fused_optimizer and all(p.dtype == torch.float32 for p in self.params)
        self._flat = {}
        if self.fused:
            # fp32 master weights: parameters, gradients, Adam moments and the bf16 shadow weights of
            # each weight-decay group live in flat buffers (every tensor padded to 8 elements: the fp32
            # segments and the bf16 shadows -- TMA operands -- all start 16-byte aligned),
            # so one optimizer step is two launches of clipa_adamw_step and the data-parallel
            # gradient reduction is one all-reduce per group.
            self._groups = []
            self._param_slot = {}          # id(param) -> (group index, offset, numel) in the flat buffers
            for ps, group_wd in ((gain, 0.0), (rest, wd)):
                if not ps:
                    continue
                sizes = [(p.numel() + 7) // 8 * 8 for p in ps]
                total = sum(sizes)
                fp = torch.zeros(total, dtype=torch.float32, device=dev)
                fg = torch.zeros(total, dtype=torch.float32, device=dev)
                fb = torch.zeros(total, dtype=torch.bfloat16, device=dev)
                off = 0
                for p, sz in zip(ps, sizes):
                    n = p.numel()
                    fp[off:off + n].copy_(p.data.reshape(-1))
                    p.data = fp[off:off + n].view_as(p)
                    p.grad = fg[off:off + n].view_as(p)
                    p._clipa_direct_grad = True   # gradient kernels accumulate straight into the flat buffer
                    shadow = fb[off:off + n].view_as(p)
                    shadow.copy_(p.data)
                    p._clipa_bf16 = (p._version, shadow)
                    self._param_slot[id(p)] = (len(self._groups), off, n)
                    off += sz
                self._groups.append(dict(p=fp, g=fg, m=torch.zeros_like(fp), v=torch.zeros_like(fp), b=fb, wd=group_wd,
                                         params=ps, sizes=sizes))
                self._flat[len(self._groups)] = fg
            self.optimizer = None
        else:
            # pure-bf16 parameters (or fused_optimizer=False): flat gradient storage per dtype + torch AdamW
            for dt in {p.dtype for p in self.params}:
                ps = [p for p in self.params if p.dtype == dt]
                flat = torch.zeros(sum(p.numel() for p in ps), dtype=dt, device=dev)
                off = 0
                for p in ps:
                    p.grad = flat[off:off + p.numel()].view_as(p)
                    off += p.numel()
                self._flat[dt] = flat
            self.optimizer = torch.optim.AdamW(
                [{"params": gain, "weight_decay": 0.}, {"params": rest, "weight_decay": wd}],
                lr=lr, betas=betas, eps=eps, fused=dev.type == "cuda")

    # ---- gradient all-reduce overlapped with backward (the role of DDP's buckets, training/main.py:299) ----------
    # Gradients land directly in the flat buffers, and a tower's blocks finish their backward last-to-first, so a
    # bucket of `allreduce_bucket_blocks` consecutive blocks is one contiguous range per weight-decay group that is
    # complete as soon as d(input of its first block) has been computed.  Transformer.forward puts a tensor hook
    # there; the hook launches the NCCL all-reduce of the range on a side stream while the earlier blocks are
    # still back-propagating.  What no bucket covers (embeddings, projections, ln_pre/post, logit_scale) is reduced
    # after backward.  Every rank builds the same graph, so the hooks fire -- and the collectives are issued -- in
    # the same order everywhere.
    def _setup_overlap(self):
        self._buckets = {}
        towers = []
        for mod in self.model.modules():
            if type(mod).__name__ == "Transformer" and hasattr(mod, "resblocks"):
                towers.append(mod)
        for tower in towers:
            blocks = list(tower.resblocks)
            for b0 in range(0, len(blocks), self._bucket_blocks):
                ranges = {}
                for blk in blocks[b0:b0 + self._bucket_blocks]:
                    for p in blk.parameters():
                        slot = self._param_slot.get(id(p))
                        if slot is None:
                            continue
                        gi, off, n = slot
                        lo, hi = ranges.get(gi, (1 << 62, 0))
                        ranges[gi] = (min(lo, off), max(hi, off + n))
                if ranges:
                    self._buckets[(id(tower), b0)] = [(gi, lo, hi) for gi, (lo, hi) in sorted(ranges.items())]
            tower.grad_bucket_blocks = self._bucket_blocks
            tower.grad_ready_callback = self._on_blocks_ready
        self._side = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

    def _on_blocks_ready(self, tower, first_block: int):
        """Called from the autograd engine when every block >= first_block of `tower` has its gradients."""
        if not self._armed:
            return
        ranges = self._buckets.get((id(tower), first_block))
        if not ranges:
            return
        if self._side is not None:
            ev = torch.cuda.Event()
            ev.record()                               # the wgrads / reductions of these blocks are in flight before it
            self._side.wait_event(ev)
        with (torch.cuda.stream(self._side) if self._side is not None else contextlib.nullcontext()):
            for gi, lo, hi in ranges:
                self._works.append(dist.all_reduce(self._groups[gi]["g"][lo:hi], op=dist.ReduceOp.SUM, async_op=True))
                self._done_ranges.append((gi, lo, hi))

    def _finish_overlapped_allreduce(self):
        """All-reduce whatever the buckets did not cover, then make the compute stream wait for all of it."""
        for gi, grp in enumerate(self._groups):
            done = sorted((lo, hi) for g, lo, hi in self._done_ranges if g == gi)
            pos, total = 0, grp["g"].numel()
            for lo, hi in done + [(total, total)]:
                if lo > pos:
                    self._works.append(dist.all_reduce(grp["g"][pos:lo], op=dist.ReduceOp.SUM, async_op=True))
                pos = max(pos, hi)
        for w in self._works:
            w.wait()
        self._works, self._done_ranges = [], []
        self._armed = False

    # -------------------------------------------------------------------------------------
    def preprocess(self, images: torch.Tensor) -> torch.Tensor:
        """uint8 [B,3,H,W] (host or device) -> normalised bf16 on device (train.py:191-197)."""
        images = images.to(self.device, non_blocking=True)
        if images.dtype == torch.uint8:
            if images.is_cuda and images.dim() == 4 and images.shape[1] == 3:
                from . import ops
                return ops.preprocess_u8(images.contiguous(), self._mean_host, self._std_host)   # one pass, own kernel
            images = (images.float().div_(255.0) - self._mean) * self._inv_std
        return images.to(torch.bfloat16)

    _copy_stream = None
    _ready = ()
    stage_on_copy_stream = True

    def _stage_host_images(self, images: torch.Tensor) -> torch.Tensor:
        """Host uint8 batch -> normalised bf16 on the device, one micro-batch at a time on a COPY stream: the H2D
        transfer and the normalise kernel of chunk c overlap with whatever the compute stream is doing (the text
        tower of the first chunk, the towers of earlier chunks).  Each chunk gets a ready event; the chunk's
        slice carries it (`_clipa_ready`) and CLIP.encode_image makes the compute stream wait on it -- after the
        text tower, which needs no image."""
        from . import ops
        B, mb = images.shape[0], self.micro_batch
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        side, main = self._copy_stream, torch.cuda.current_stream()
        # persistent staging buffers (uint8 landing area + normalised bf16 batch): no allocator traffic per step and no
        # cross-stream block reuse to police -- the copy stream first waits for everything queued on the compute
        # stream (the previous step), then overwrites them
        key = tuple(images.shape)
        if getattr(self, "_stage_key", None) != key:
            self._stage_u8 = torch.empty(images.shape, dtype=torch.uint8, device=self.device)
            self._stage_bf16 = torch.empty(images.shape, dtype=torch.bfloat16, device=self.device)
            self._stage_key = key
        u8, out = self._stage_u8, self._stage_bf16
        side.wait_stream(main)
        ready = []
        with torch.cuda.stream(side):
            for s in range(0, B, mb):
                e = min(B, s + mb)
                u8[s:e].copy_(images[s:e], non_blocking=True)
                ops.preprocess_u8(u8[s:e], self._mean_host, self._std_host, out=out[s:e])
                ev = torch.cuda.Event()
                ev.record(side)
                ready.append((s, e, ev))
        self._ready = ready
        return out

    def _chunk(self, images: torch.Tensor, s: int, e: int) -> torch.Tensor:
        x = images[s:e]
        evs = [ev for (a, b, ev) in self._ready if a < e and b > s]
        if evs:
            x._clipa_ready = evs
        return x

    def zero_grad(self):
        for flat in self._flat.values():
            flat.zero_()

    def _allreduce_grads(self, average: bool = True):
        if self.world_size > 1:
            if self._armed:                 # buckets were launched during backward: reduce the rest and wait
                self._finish_overlapped_allreduce()
            else:
                for flat in self._flat.values():
                    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            if average:
                for flat in self._flat.values():
                    flat.div_(self.world_size)

    def optimizer_step(self, grad_scale: float = 1.0, grad_scale_dev: Optional[torch.Tensor] = None):
        """AdamW update; with the fused kernel the 1/world_size averaging, the bf16 shadow refresh and the
        clearing of the gradient buffers ride along in the same pass."""
        if not self.fused:
            self.optimizer.step()
            # torch's fused/foreach optimizers may update parameters without bumping their version
            # counter, so the bf16 shadow weights are refreshed explicitly rather than lazily
            with torch.no_grad():
                for p in self.params:
                    cache = getattr(p, "_clipa_bf16", None)
                    if cache is not None:
                        cache[1].copy_(p.detach())
                        p._clipa_bf16 = (p._version, cache[1])
            return
        from . import ops
        self.step_count += 1
        for grp in self._groups:
            ops.adamw_step(grp["p"], grp["g"], grp["m"], grp["v"], grp["b"], lr=self.lr, beta1=self.betas[0],
                           beta2=self.betas[1], eps=self.eps, weight_decay=grp["wd"], step=self.step_count,
                           grad_scale=grad_scale, grad_scale_dev=grad_scale_dev, zero_grad=True)

    # ---- checkpoint / resume (training/main.py:352-369 saves optimizer.state_dict(); --resume restores it) ----
    def state_dict(self) -> dict:
        """torch.optim.AdamW-compatible optimizer state (same param order as main.py:318-326: the no-decay group
        first, then the decayed one), so checkpoints move between the reference's optimizer and this step."""
        if not self.fused:
            return self.optimizer.state_dict()
        state, groups, idx = {}, [], 0
        for grp in self._groups:
            ids, off = [], 0
            for p, sz in zip(grp["params"], grp["sizes"]):
                n = p.numel()
                state[idx] = {"step": torch.tensor(float(self.step_count)),
                              "exp_avg": grp["m"][off:off + n].view_as(p).clone(),
                              "exp_avg_sq": grp["v"][off:off + n].view_as(p).clone()}
                ids.append(idx)
                idx += 1
                off += sz
            groups.append({"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": grp["wd"],
                           "amsgrad": False, "maximize": False, "params": ids})
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd: dict) -> None:
        if not self.fused:
            self.optimizer.load_state_dict(sd)
            return
        steps = set()
        for grp, g_sd in zip(self._groups, sd["param_groups"]):
            assert len(g_sd["params"]) == len(grp["params"]), "optimizer state does not match the parameter groups"
            off = 0
            for p, sz, pid in zip(grp["params"], grp["sizes"], g_sd["params"]):
                st = sd["state"].get(pid)
                n = p.numel()
                if st is not None:
                    grp["m"][off:off + n].copy_(st["exp_avg"].reshape(-1))
                    grp["v"][off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.add(int(float(st["step"])))
                off += sz
        if sd["param_groups"]:
            self.lr = sd["param_groups"][0].get("lr", self.lr)
        assert len(steps) <= 1, f"per-parameter step counts differ: {sorted(steps)}"
        self.step_count = steps.pop() if steps else 0

    def forward_backward(self, images: torch.Tensor, texts: torch.Tensor) -> torch.Tensor:
        model = self.model
        B = images.shape[0]
        mb = self.micro_batch
        overlap = self.overlap and self.fused
        if overlap and not hasattr(self, "_buckets"):
            self._setup_overlap()
        # HBM the contrastive head will need at the end of the forward pass (gathered features, one softmax-gradient
        # block per direction, fp32 feature gradients): the towers' activation policies leave room for it
        E = getattr(getattr(model, "visual", None), "output_dim", 1024)
        bg = B * self.world_size
        model.head_reserve_bytes = (2 * bg * E * 2 + 2 * min(min(B, mb) * bg * 2, 512 << 20) + 2 * bg * E * 4
                                    + 2 * B * E * 4 + (1 << 30))
        if B <= mb:
            out = model(self._chunk(images, 0, B), texts)
            feats_i, feats_t, scale = self._unpack(out)
            loss = self.loss_fn(feats_i, feats_t, scale)
            self._armed = overlap            # this backward completes the gradients: buckets may go out as they finish
            loss.backward()
            return loss.detach()
        chunks = [(s, min(B, s + mb)) for s in range(0, B, mb)]
        # GradCache (train.py:216-256): features of every chunk without autograd, the loss once on the
        # whole batch, then each chunk again with autograd.  The LAST chunk of the feature pass keeps
        # its graph -- only one chunk's activations are alive at a time either way -- so its second
        # forward is saved: N chunks cost 2N-1 forwards instead of 2N (same numbers, same peak memory).
        (ls, le) = chunks[-1]
        # Stochastic layers (PatchDropout draws fresh scores per call, transformer.py:76-82) must keep the SAME
        # tokens in a chunk's second forward, or the cached d(features) would be applied to features they were
        # not computed for: snapshot the device RNG before each no-grad forward and restore it for the re-run
        # (GradCache's RandContext).  The reference recomputes the loss from the re-run features instead.
        stochastic = model.training and any(getattr(m, "prob", 0.) > 0. for m in model.modules()
                                            if type(m).__name__ == "PatchDropout")
        rng = []
        with torch.no_grad():
            fi, ft = [], []
            for s, e in chunks[:-1]:
                if stochastic:
                    rng.append(torch.cuda.get_rng_state(self.device))
                a, b, _ = self._unpack(model(self._chunk(images, s, e), texts[s:e]))
                fi.append(a)
                ft.append(b)
        a_last, b_last, _ = self._unpack(model(self._chunk(images, ls, le), texts[ls:le]))
        fi = torch.cat(fi + [a_last.detach()]).requires_grad_(True)
        ft = torch.cat(ft + [b_last.detach()]).requires_grad_(True)
        loss = self.loss_fn(fi, ft, model.logit_scale.exp())
        loss.backward()
        if self.reference_accum_logit_scale and model.logit_scale.grad is not None:
            model.logit_scale.grad.mul_(len(chunks))
        torch.autograd.backward([a_last, b_last], [fi.grad[ls:le], ft.grad[ls:le]])
        del a_last, b_last
        if stochastic:
            after = torch.cuda.get_rng_state(self.device)
        for c, (s, e) in enumerate(chunks[:-1]):
            if stochastic:
                torch.cuda.set_rng_state(rng[c], self.device)
            a, b, _ = self._unpack(model(images[s:e], texts[s:e]))
            self._armed = overlap and c == len(chunks) - 2      # the LAST backward completes the accumulated gradients
            torch.autograd.backward([a, b], [fi.grad[s:e], ft.grad[s:e]])
        if stochastic:
            torch.cuda.set_rng_state(after, self.device)
        return loss.detach()

    @staticmethod
    def _unpack(out):
        if isinstance(out, dict):
            return out["image_features"], out["text_features"], out["logit_scale"]
        return out

    def step(self, images: torch.Tensor, texts: torch.Tensor) -> torch.Tensor:
        """images: uint8 or float [B_local,3,H,W]; texts: int64 [B_local, ctx].  Returns the loss
        (device scalar, no host sync)."""
        self._ready = ()
        if (not images.is_cuda and images.dtype == torch.uint8 and self.device.type == "cuda" and images.dim() == 4
                and images.shape[1] == 3 and self.stage_on_copy_stream):
            texts = texts.to(self.device, non_blocking=True)      # small; first, the text tower starts with it
            images = self._stage_host_images(images)
        else:
            images = self.preprocess(images)
            texts = texts.to(self.device, non_blocking=True)
        if not (self.fused and self.step_count > 0):
            self.zero_grad()              # the fused optimizer kernel leaves the buffers cleared
        loss = self.forward_backward(images, texts)
        self._allreduce_grads(average=not self.fused)
        scale = 1.0 / self.world_size if self.fused else 1.0
        clip = None
        if self.grad_clip_norm is not None:
            if self.fused:
                # clip_grad_norm_ (train.py:277-283): factor min(1, max_norm / (||g|| + 1e-6)) with g the averaged
                # gradient; computed on the device and handed to the optimizer kernel as a pointer (no .item())
                total = torch.sqrt(sum(torch.linalg.vector_norm(f) ** 2 for f in self._flat.values())) * scale
                clip = torch.clamp(self.grad_clip_norm / (total + 1e-6), max=1.0).reshape(1).float()
            else:
                torch.nn.utils.clip_grad_norm_(self.params, self.grad_clip_norm, norm_type=2.0)
        self.optimizer_step(grad_scale=scale, grad_scale_dev=clip)
        with torch.no_grad():
            self.model.logit_scale.clamp_(0, math.log(100))
        return loss
