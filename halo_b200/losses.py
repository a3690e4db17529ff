"""Losses on the head's logits, fused with the up-sampling in front of them and its adjoint -- SURVEY section 8f row 4.

The reference learner up-samples the low-resolution logits to the crop size inside the classifier
(core/models/classifier.py:556-557), then takes softmax, CrossEntropyLoss(ignore_index=255) on the labelled pixels and the
negative-learning loss (core/train_learners.py:343-356; core/loss/negative_learning_loss.py:6-16), and back-propagates
through all of it.  `fused_seg_loss` does the same from the LOW-resolution logits in two kernels (`halo_seg_loss`): nothing
of size (N,O,H,W) is materialised, forward or backward, and the gradient is bitwise reproducible (a gather, unlike torch's
atomic scatter in upsample_bilinear2d_backward).  A maintainer who wants it replaces, in the learner,

    tgt_out = self.forward(tgt_input)[0]; predict = softmax(tgt_out); loss = CE(tgt_out, tgt_mask) + neg(predict) * w

by

    logits_lr = self.classifier(self.feature_extractor(tgt_input))[0]            # size=None: no up-sampling in the head
    loss, loss_sup, negative_loss = fused_seg_loss(logits_lr, tgt_mask, tgt_input.shape[-2:], w)

(INTEGRATION.md).  There is no CPU path.
"""
import torch

from . import _native as nat


class _FusedSegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_lr, labels, size, neg_weight, threshold):
        lib = nat.load()
        nat.require_cuda(logits_lr, "logits")
        x = logits_lr.detach().float().contiguous()
        N, O, h, w = x.shape
        H, W = int(size[0]), int(size[1])
        lab8 = None
        if labels is not None:
            nat.require_cuda(labels, "labels")
            lab8 = labels.to(torch.uint8).reshape(N, H, W).contiguous()
        losses = torch.empty((4,), dtype=torch.float32, device=x.device)
        need_grad = logits_lr.requires_grad
        dl = torch.empty_like(x) if need_grad else None
        ws = nat.workspace.get(x.device, "seg_loss", lib.halo_seg_loss_workspace_bytes())
        with torch.cuda.device(x.device):
            rc = lib.halo_seg_loss(nat.ptr(x), nat.ptr(lab8), float(neg_weight), float(threshold), nat.ptr(losses), nat.ptr(dl),
                                   N, O, h, w, H, W, nat.ptr(ws), ws.numel(), nat.stream_of(x))
        nat.check(rc, "halo_seg_loss")
        ctx.save_for_backward(dl)
        ctx.in_dtype = logits_lr.dtype
        ctx.mark_non_differentiable(losses)
        return losses[0].clone(), losses

    @staticmethod
    def backward(ctx, g_total, _g_losses):
        (dl,) = ctx.saved_tensors
        return (dl * g_total).to(ctx.in_dtype), None, None, None, None


def fused_seg_loss(logits_lr, labels, size, neg_weight=1.0, threshold=0.05):
    """CrossEntropy(ignore 255) + negative-learning loss of the bilinearly up-sampled logits.

    logits_lr (N,O,h,w) CUDA float tensor (requires_grad for training); labels (N,H,W) integer tensor with 255 = unlabelled,
    or None (no supervised term); size = (H,W) of the up-sampled logits; neg_weight = SOLVER.NEGATIVE_LOSS (0 disables).
    Returns (loss, loss_sup, negative_loss) as 0-d device tensors -- `loss` carries the gradient; the other two are the
    values the reference logs (train_learners.py:347,353).  No host synchronisation."""
    total, losses = _FusedSegLoss.apply(logits_lr, labels, tuple(size), float(neg_weight), float(threshold))
    return total, losses[1], losses[2]
