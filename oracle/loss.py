"""Oracle for the training losses on the head's logits (ORACLE -- test infrastructure, CPU only).

Restates, with torch CPU ops exactly as the reference chains them:
  * the up-sampling of the logits inside the classifier     core/models/classifier.py:556-557 (and :376-377)
  * softmax / CrossEntropyLoss(ignore_index=255)             core/train_learners.py:343-348 (criterion built at :47)
  * NegativeLearningLoss(threshold=0.05)                     core/loss/negative_learning_loss.py:6-16, weighted at
                                                             core/train_learners.py:351-353
and the autograd backward the learner takes through all of it (manual_backward, :362).
"""
import torch
import torch.nn.functional as F


def negative_learning_loss(predict, threshold=0.05):
    """negative_learning_loss.py:11-16 -- the mask is detached; 0/0 when nothing is below the threshold."""
    mask = (predict < threshold).detach()
    item = -1 * mask * torch.log(1 - predict + 1e-6)
    return torch.sum(item) / torch.sum(mask)


def seg_loss(logits_lr, labels, size, neg_weight=1.0, threshold=0.05, dtype=torch.float64):
    """Returns (loss, loss_sup, negative_loss, dlogits_lr) with the reference's control flow: the supervised term is added
    only when some pixel is labelled (:345), the negative term only when its weight is positive (:351)."""
    x = logits_lr.detach().to(dtype).clone().requires_grad_(True)
    out = F.interpolate(x, size=tuple(size), mode="bilinear", align_corners=True)
    predict = torch.softmax(out, dim=1)
    loss = torch.zeros((), dtype=dtype)
    loss_sup = torch.zeros((), dtype=dtype)
    neg = torch.zeros((), dtype=dtype)
    if labels is not None and torch.sum(labels != 255) != 0:
        loss_sup = F.cross_entropy(out, labels.long(), ignore_index=255)
        loss = loss + loss_sup
    if neg_weight > 0:
        neg = negative_learning_loss(predict, threshold) * neg_weight
        loss = loss + neg
    (g,) = torch.autograd.grad(loss, x) if loss.requires_grad else (torch.zeros_like(x),)
    return loss.detach(), loss_sup.detach(), neg.detach(), g
