"""Drop-in for the reference's `runs/eval.py` (same import path, same function): generalized few-shot mIoU metrics.

The reference walks every point of every prediction in a Python loop (runs/eval.py:31-48: one `int()` per element, a host
sync per element when the arrays are CUDA tensors).  Here the three count vectors come from ONE joint histogram of
(gt, pred) on the GPU (gfs_joint_histogram_i32, csrc/hist.cu); the IoU arithmetic on the <= 32 x 32 integer confusion
matrix stays in Python with the reference's operation order, so the returned floats are identical.
"""
import numpy as np
import torch

from gfs3d import ops


def confusion_matrix(pred_labels_list, gt_labels_list, num_labels, device=None):
    """J[gt, pred] (int64, num_labels x num_labels, on the host) over all arrays of the two lists; raises IndexError (as the
    reference's `all_learning_order[label]` does) if a label is outside [0, num_labels)."""
    if not torch.cuda.is_available():
        raise RuntimeError("runs.eval needs a CUDA device: the metric reduction has no CPU fallback")
    device = device or torch.device("cuda", torch.cuda.current_device())
    J = torch.zeros(num_labels, num_labels, dtype=torch.int64, device=device)
    total = 0
    for pred, gt in zip(pred_labels_list, gt_labels_list):
        p = torch.as_tensor(pred).to(device)
        g = torch.as_tensor(gt).to(device)
        if p.shape != g.shape:
            raise ValueError(f"prediction {tuple(p.shape)} and ground truth {tuple(g.shape)} differ in shape")
        total += p.numel()
        ops.joint_histogram(g, p, num_labels, num_labels, out=J)
    Jh = J.cpu().numpy()
    if int(Jh.sum()) != total:
        raise IndexError("list index out of range")      # a label outside all_learning_order, as in the reference
    return Jh


def evaluate_metric_GFS(logger, pred_labels_list, gt_labels_list, test_classes, novel_classes, all_learning_order, scannet=False):
    """runs/eval.py:10-108 of the reference: class-wise IoU, mean / base / novel IoU and their harmonic mean.
    pred_labels_list / gt_labels_list: lists of (n_queries, num_points) integer arrays (numpy or torch, host or device)."""
    assert len(pred_labels_list) == len(gt_labels_list)
    logger.cprint('*****Test Classes: {0}*****'.format(test_classes))
    NUM_CLASS = len(test_classes)
    L = len(all_learning_order)
    J = confusion_matrix(pred_labels_list, gt_labels_list, L)
    gt_classes = [0 for _ in range(NUM_CLASS)]
    positive_classes = [0 for _ in range(NUM_CLASS)]
    true_positive_classes = [0 for _ in range(NUM_CLASS)]
    for lab in range(L):
        row, col = int(J[lab, :].sum()), int(J[:, lab].sum())
        if row or col:                                   # the reference only indexes the labels that occur
            idx = all_learning_order[lab]
            gt_classes[idx] += row
            positive_classes[idx] += col
            true_positive_classes[idx] += int(J[lab, lab])

    # IoU per class name, then the three means: the reference's arithmetic (integer counts -> Python float division ->
    # numpy mean over a list) in the reference's order, so the floats agree to the last bit
    iou_list = [true_positive_classes[c] / float(gt_classes[c] + positive_classes[c] - true_positive_classes[c])
                for c in range(NUM_CLASS)]
    for c, iou in enumerate(iou_list):
        logger.cprint('----- [class %d]  IoU: %f -----' % (c, iou))
    first = 1 if scannet else 0                          # ScanNet: class name 0 is skipped everywhere
    kept = range(first, NUM_CLASS)
    mean_iou = np.array([iou_list[c] for c in kept]).mean()
    base_iou = np.array([iou_list[c] for c in kept if c not in novel_classes]).mean()
    novel_iou = np.array([iou_list[c] for c in kept if c in novel_classes]).mean()
    hm = 2 * base_iou * novel_iou / (base_iou + novel_iou)
    for label, value in (('mean-iou', mean_iou), ('base-iou', base_iou), ('novel-iou', novel_iou), ('hm-iou', hm)):
        logger.cprint('{}: {}'.format(label, value))
    return mean_iou, base_iou, novel_iou, hm, np.array(iou_list[first:])
