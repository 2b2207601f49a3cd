"""Zero-shot classification on the B200 towers -- the only consumer of encode_image / encode_text outside
training (clipa_torch/training/zero_shot.py:29-117).

Same three steps as the reference: a classifier matrix from the text tower (per class: encode every prompt,
L2-normalise, average, re-normalise; zero_shot.py:29-46), logits = 100 * normalised image features @ classifier
(:73-77), top-k accuracy (:49-52).  Tokenisation is host-side string processing outside the hot path: prompts
arrive as int64 token ids (any tokenizer with the reference's output contract, [n_prompts, context_length]), or a
`tokenizer` callable is applied to the formatted strings exactly like the reference does."""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import torch

from . import ops
from .open_clip.model import _l2_normalize


def unwrap_model(model):
    return model.module if hasattr(model, "module") else model


@torch.no_grad()
def zero_shot_classifier(model, classnames: Optional[Sequence[str]] = None, templates: Optional[Sequence[Callable]] = None,
                         tokenizer: Optional[Callable] = None, *, class_token_ids: Optional[Iterable[torch.Tensor]] = None,
                         device=None) -> torch.Tensor:
    """Returns the [embed_dim, n_classes] bf16 classifier (columns = unit-norm mean prompt embeddings).
    Either (classnames, templates, tokenizer) as in zero_shot.py:29-46, or pre-tokenised `class_token_ids`
    (one [n_prompts, context_length] int64 tensor per class)."""
    m = unwrap_model(model)
    device = device or next(m.parameters()).device
    if class_token_ids is None:
        assert classnames is not None and templates is not None and tokenizer is not None
        class_token_ids = (tokenizer([t(c) for t in templates]) for c in classnames)
    weights: List[torch.Tensor] = []
    for ids in class_token_ids:
        emb = m.encode_text(ids.to(device))                       # [n_prompts, E], un-normalised (zero_shot.py:36)
        emb = _l2_normalize(emb).float().mean(dim=0)
        weights.append(emb / emb.norm())
    return torch.stack(weights, dim=1).to(torch.bfloat16).contiguous()


def accuracy(output: torch.Tensor, target: torch.Tensor, topk: Tuple[int, ...] = (1,)) -> List[float]:
    """zero_shot.py:49-52."""
    pred = output.topk(max(topk), 1, True, True)[1].t()
    correct = pred.eq(target.view(1, -1).expand_as(pred))
    return [float(correct[:k].reshape(-1).float().sum().item()) for k in topk]


@torch.no_grad()
def classify(model, classifier: torch.Tensor, images: torch.Tensor) -> torch.Tensor:
    """logits [batch, n_classes] fp32 = 100 * normalize(encode_image(images)) @ classifier (zero_shot.py:73-77);
    the product runs on the tcgen05 GEMM (classifier columns read in place as the MN-major B operand)."""
    m = unwrap_model(model)
    feats = _l2_normalize(m.encode_image(images)).contiguous()
    n_cls = classifier.shape[1]
    if n_cls % 8 or feats.shape[1] % 8:       # TMA pitch: tiny class counts go through torch
        return 100.0 * feats.float() @ classifier.float()
    logits = torch.empty(feats.shape[0], n_cls, dtype=torch.float32, device=feats.device)
    ops.gemm(feats, classifier.t(), logits, alpha=100.0)
    return logits


@torch.no_grad()
def run(model, classifier: torch.Tensor, dataloader, device=None, preprocess: Optional[Callable] = None) -> Tuple[float, float]:
    """zero_shot.py:55-93: top-1 / top-5 over a dataloader of (images, target); `preprocess` is e.g.
    TrainStep.preprocess for uint8 batches (--to-float-on-device)."""
    m = unwrap_model(model)
    device = device or next(m.parameters()).device
    top1 = top5 = n = 0.0
    for images, target in dataloader:
        if isinstance(images, (list, tuple)):
            images = images[0]
        images = images.to(device)
        if preprocess is not None:
            images = preprocess(images)
        logits = classify(model, classifier, images)
        a1, a5 = accuracy(logits, target.to(device), topk=(1, min(5, logits.shape[1])))
        top1 += a1
        top5 += a5
        n += images.size(0)
    return top1 / n, top5 / n
