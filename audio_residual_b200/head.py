"""Classification head on the joint embedding, in libard_b200.so (SURVEY §8b `ard_head_fwd/bwd`).

* `head_logits(emb, W, bias=None)`  = emb @ W.T (+ bias): the zero-shot similarities `audio_embeds @ text_embeddings.T`
  (src/training.py:28, src/evaluation.py:98) and the linear probe `nn.Linear(512, n_classes)` (src/linear.py:23-32).
* `cross_entropy(logits, labels)`   = nn.CrossEntropyLoss() (mean reduction; src/training.py:29, src/linear.py:43).
Both are autograd Functions whose forward and backward are kernels of the library (no torch.matmul / cuBLAS).
* `eval_metrics(scores, targets, k)` = top-1 / top-k hit counts, confusion matrix and argmax predictions as device
  reductions (the numbers of visualize_eval_metrics, src/evaluation.py:159-177).
"""
import torch

from . import lib as L


def _f32c(t, dev):
    return t.detach().to(dev, torch.float32).contiguous()


class _LogitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, W, bias):
        dev = emb.device
        if dev.type != "cuda":
            raise RuntimeError("audio_residual_b200 head runs on CUDA only (no CPU fallback)")
        e, w = _f32c(emb, dev), _f32c(W, dev)
        b = _f32c(bias, dev) if bias is not None else None
        B, J = e.shape
        N = w.shape[0]
        if w.shape[1] != J:
            raise ValueError(f"head: embedding width {J} does not match weight {tuple(w.shape)}")
        out = torch.empty((B, N), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            L.check(L.load().ard_head_forward(L.ptr(e), L.ptr(w), L.ptr(b), B, N, J, L.ptr(out), L.stream_ptr()))
        ctx.save_for_backward(e, w)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        e, w = ctx.saved_tensors
        dev = e.device
        g = _f32c(g, dev)
        B, J = e.shape
        N = w.shape[0]
        need_e, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]
        d_e = torch.empty_like(e) if need_e else None
        d_w = torch.empty_like(w) if need_w else None
        d_b = torch.empty(N, device=dev, dtype=torch.float32) if need_b else None
        with torch.cuda.device(dev):
            L.check(L.load().ard_head_backward(L.ptr(g), L.ptr(e), L.ptr(w), B, N, J, L.ptr(d_e), L.ptr(d_w), L.ptr(d_b), L.stream_ptr()))
        return d_e, d_w, d_b


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        dev = logits.device
        z = _f32c(logits, dev)
        lab = labels.detach().to(dev, torch.int64).contiguous()
        B, N = z.shape
        if lab.shape != (B,):
            raise ValueError(f"cross_entropy: labels {tuple(lab.shape)} do not match logits {tuple(z.shape)}")
        if int(lab.min()) < 0 or int(lab.max()) >= N:
            raise IndexError("Target out of bounds")          # what nn.CrossEntropyLoss raises
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        dz = torch.empty_like(z)
        with torch.cuda.device(dev):
            L.check(L.load().ard_ce_forward(L.ptr(z), L.ptr(lab), B, N, L.ptr(loss), L.ptr(dz), L.stream_ptr()))
        ctx.save_for_backward(dz)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None


def head_logits(emb, W, bias=None):
    return _LogitsFn.apply(emb, W, bias)


def cross_entropy(logits, labels):
    return _CEFn.apply(logits, labels)


def apply_criterion(criterion, logits, labels):
    """The drivers take `criterion` as an argument (src/training.py:12, src/linear.py:35). A default nn.CrossEntropyLoss - what
    every caller in the reference passes - runs in the library; any other callable is the caller's own code and is called as is."""
    ce = torch.nn.CrossEntropyLoss
    if (type(criterion) is ce and criterion.weight is None and criterion.reduction == "mean" and criterion.label_smoothing == 0.0
            and criterion.ignore_index == -100 and logits.is_cuda):
        return cross_entropy(logits, labels)
    return criterion(logits, labels)


def eval_metrics(scores, targets, k=5, n_classes=None):
    """scores [n, C] (tensor or ndarray), targets [n] -> dict(top1, topk, n, confusion [C, C] int64 (rows = true), predictions [n])."""
    dev = scores.device if torch.is_tensor(scores) and scores.is_cuda else torch.device("cuda", torch.cuda.current_device())
    sc = torch.as_tensor(scores).to(dev, torch.float32).contiguous()
    tg = torch.as_tensor(targets).to(dev, torch.int64).contiguous()
    n, C = sc.shape
    if n_classes is not None and n_classes != C:
        raise ValueError(f"scores have {C} classes, expected {n_classes}")
    k = min(k, C)                                             # src/evaluation.py:160 k_eff
    counts = torch.zeros(2, device=dev, dtype=torch.int64)
    cm = torch.zeros((C, C), device=dev, dtype=torch.int64)
    preds = torch.empty(n, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        L.check(L.load().ard_eval_metrics(L.ptr(sc), L.ptr(tg), n, C, k, L.ptr(counts), L.ptr(cm), L.ptr(preds), L.stream_ptr()))
    c = counts.cpu()
    return {"top1": int(c[0]) / max(n, 1), "topk": int(c[1]) / max(n, 1), "k": k, "n": n, "confusion": cm.cpu().numpy(),
            "predictions": preds.cpu().numpy()}
