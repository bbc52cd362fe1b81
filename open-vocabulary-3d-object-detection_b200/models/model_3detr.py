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

    @torch.no_grad()
    def forward(self, visual_embeds, want_logits=False):
        """visual_embeds [..., 640] -> dict(sem_cls_logits, sem_cls_prob, objectness_prob)."""
        lg, pr, ob = clip_logits(visual_embeds, self.weight, self.l2norm, self.logit_scale, want_logits=want_logits)
        return {"sem_cls_logits": lg, "sem_cls_prob": pr, "objectness_prob": ob}
