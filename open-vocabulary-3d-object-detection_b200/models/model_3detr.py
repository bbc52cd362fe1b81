"""Drop-in for the open-vocabulary classification head of the reference's
``models/model_3detr.py``: the frozen CLIP-text ``sem_cls_head`` (:151-154, :238)
and ``BoxProcessor.compute_objectness_and_cls_prob`` (:58-62), fused into one
tcgen05/TMA kernel (csrc/clip_logits.cu).

``ClipTextClassifier`` is what a maintainer swaps in for the
``nn.Linear(640, C+1, bias=False)`` + softmax pair: logits are produced on the
5th-gen tensor cores from bf16 operands with fp32 accumulation in TMEM, the row
softmax is finished across the N-tiles of a thread-block cluster and the
probabilities are written once, in bf16 (fp32 logits optional)."""
import torch
import torch.nn as nn

from .. import _capi as C


def clip_logits(x, text, l2norm=False, scale=1.0, want_logits=False, want_prob=True):
    """x [..., K], text [N, K] (any float dtype; converted to bf16).  Returns
    (logits fp32 [..., N] or None, sem_cls_prob bf16 [..., N-1] (a view of the
    padded softmax buffer, like the reference's prob[..., :-1]), objectness fp32 [...])."""
    C.require_cuda(x, text)
    dev = x.device
    lead = x.shape[:-1]
    K = x.shape[-1]
    xb = x.detach().reshape(-1, K).to(torch.bfloat16).contiguous()
    tb = text.detach().to(device=dev, dtype=torch.bfloat16).contiguous()
    M, N = xb.shape[0], tb.shape[0]
    assert tb.shape[1] == K
    ldp = (N + 7) // 8 * 8
    logits = torch.empty((M, N), dtype=torch.float32, device=dev) if want_logits else None
    prob = torch.empty((M, ldp), dtype=torch.bfloat16, device=dev) if want_prob else None
    obj = torch.empty((M,), dtype=torch.float32, device=dev)
    flags = C.LOGITS_L2NORM if l2norm else 0
    with torch.cuda.device(dev):
        C.check(C.lib().ovdet_clip_logits_bf16(C.ptr(xb), C.ptr(tb), M, K, N, flags, float(scale), C.ptr(logits), N,
                                               C.ptr(prob), ldp, C.ptr(obj), C.stream(dev)))
    lg = None if logits is None else logits.reshape(*lead, N)
    pr = None if prob is None else prob[:, :N - 1].reshape(*lead, N - 1) if len(lead) == 1 else (
        None if prob is None else prob.reshape(*lead, ldp)[..., :N - 1])
    return lg, pr, obj.reshape(lead)


class _ClipLogitsFn(torch.autograd.Function):
    """Differentiable form of the classification head: the forward is the fused kernel (fp32 logits + bf16 probabilities
    + objectness), the backward is the one the frozen ``nn.Linear`` of the reference has w.r.t. its input,
    dX = scale * dLogits @ T (the text matrix is frozen, models/model_3detr.py:151-154).  Probabilities and objectness
    are outputs of convenience (matcher / evaluation inputs) and carry no gradient -- the reference's losses
    differentiate ``sem_cls_logits`` (criterion.py: loss_sem_cls)."""

    @staticmethod
    def forward(ctx, x, text, scale):
        lg, pr, ob = clip_logits(x, text, False, scale, want_logits=True, want_prob=True)
        ctx.save_for_backward(text)
        ctx.scale, ctx.xdtype = float(scale), x.dtype
        ctx.mark_non_differentiable(pr, ob)
        return lg, pr, ob

    @staticmethod
    def backward(ctx, g_lg, g_pr, g_ob):
        (text,) = ctx.saved_tensors
        n = g_lg.shape[-1]
        gx = (g_lg.reshape(-1, n).to(torch.float32) @ text.to(torch.float32)) * ctx.scale
        return gx.reshape(*g_lg.shape[:-1], text.shape[1]).to(ctx.xdtype), None, None


class BoxProcessor(object):
    """The classification part of models/model_3detr.py:19-69."""

    def __init__(self, dataset_config):
        self.dataset_config = dataset_config

    def compute_objectness_and_cls_prob(self, cls_logits):
        """:58-62 on already materialised logits (kept for interface parity; a plain
        softmax -- the fused path is ClipTextClassifier)."""
        assert cls_logits.shape[-1] == self.dataset_config.num_semcls + 1
        cls_prob = torch.nn.functional.softmax(cls_logits, dim=-1)
        objectness_prob = 1 - cls_prob[..., -1]
        return cls_prob[..., :-1], objectness_prob


class ClipTextClassifier(nn.Module):
    """sem_cls_head (:151-154) + compute_objectness_and_cls_prob (:58-62) in one kernel."""

    def __init__(self, text_embedding, l2norm=False, logit_scale=1.0):
        super().__init__()
        self.register_buffer("weight", text_embedding.detach().to(torch.bfloat16).contiguous(), persistent=False)
        self.l2norm = l2norm
        self.logit_scale = logit_scale

    def forward(self, visual_embeds, want_logits=False, prob_dtype=None):
        """visual_embeds [..., 640] -> dict(sem_cls_logits, sem_cls_prob, objectness_prob).

        Only the weight is frozen in the reference: when ``visual_embeds`` requires grad (training), ``sem_cls_logits``
        is returned in fp32 WITH its graph (dX = dLogits @ T), so swapping this module in does not cut the gradient
        into ``visual_embed_head``.  ``sem_cls_prob`` comes from the kernel in bf16 (8 bits of mantissa: scores
        prob * objectness are then quantised to ~0.4 % relative, which can tie neighbouring detections in AP ranking
        and the matcher's class cost); pass ``prob_dtype=torch.float32`` to get the probabilities recomputed in fp32
        from the fp32 logits instead (one extra softmax kernel), as the reference has them."""
        if torch.is_grad_enabled() and visual_embeds.requires_grad:
            if self.l2norm:
                raise C.OvdetError("the differentiable path covers the reference's head (no normalisation, models/model_3detr.py:237-238)")
            lg, pr, ob = _ClipLogitsFn.apply(visual_embeds, self.weight, float(self.logit_scale))
        else:
            with torch.no_grad():
                lg, pr, ob = clip_logits(visual_embeds, self.weight, self.l2norm, self.logit_scale,
                                         want_logits=want_logits or prob_dtype is torch.float32)
        if prob_dtype is torch.float32:
            p32 = torch.softmax(lg.detach(), dim=-1)
            pr, ob = p32[..., :-1], 1 - p32[..., -1]
        elif prob_dtype is not None and pr is not None:
            pr = pr.to(prob_dtype)
        return {"sem_cls_logits": lg if (want_logits or lg is not None and lg.requires_grad) else None, "sem_cls_prob": pr, "objectness_prob": ob}
