"""(Un)normalise regression offsets in place (reference modules/compute_groundtruth/compute_offsets.py:6-18)."""


def normalize_gt_offsets(gt_offsets_img, offset_mu, offset_sigma):
    for d in (0, 1):
        gt_offsets_img[..., d] = (gt_offsets_img[..., d] - offset_mu[d]) / offset_sigma[d]
    return gt_offsets_img


def unnormalize_gt_offsets(offsets_img, offset_mu, offset_sigma):
    for d in (0, 1):
        offsets_img[..., d] = offsets_img[..., d] * offset_sigma[d] + offset_mu[d]
    return offsets_img
