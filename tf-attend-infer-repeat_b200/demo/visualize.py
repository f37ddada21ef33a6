"""Attention-window visualisation of the reference's image summary (air_model.py:130-158, 211-267, 627-640).

``visualize_reconstructions`` returns what the reference hands to ``tf.summary.image("reconstruction", ...)``:
per image, the 2x-enlarged original and the 2x-enlarged reconstruction side by side (a 4-pixel white stripe between
them), each with the attention windows of the executed steps drawn as red / green / blue frames.  The frames are the
border of the window, mapped onto the enlarged canvas by the write-back Spatial Transformer with the step's inverse
matrix (``rec_st_back``) -- on the device through the same ``air_st_forward`` kernel as the model (``transformer``
argument; tests on the CPU pass the oracle's).  The few elementwise steps around it are torch ops on the tensors'
device, one op per TF op so the result is bit-identical to the reference graph's (tests/test_reference_graph.py).
"""
from __future__ import annotations

import torch


def resize_bilinear(images, size):
    """tf.image.resize_images(..., method=BILINEAR, align_corners=False) of TF 1.x on NHWC input: source coordinate
    = index * (in / out), neighbours clamped at the far edge, x interpolated before y."""
    B, H, W, C = images.shape
    oh, ow = size

    def axis(n_in, n_out):
        scale = torch.tensor(n_in, dtype=torch.float32) / torch.tensor(n_out, dtype=torch.float32)
        src = torch.arange(n_out, dtype=torch.float32) * scale
        lo = torch.floor(src).long()
        return (lo.to(images.device), torch.clamp(lo + 1, max=n_in - 1).to(images.device),
                (src - lo.float()).to(images.device, images.dtype))
    y0, y1, ly = axis(H, oh)
    x0, x1, lx = axis(W, ow)
    lx, ly = lx.reshape(1, 1, ow, 1), ly.reshape(1, oh, 1, 1)
    rows0, rows1 = images[:, y0], images[:, y1]
    tl, tr, bl, br = rows0[:, :, x0], rows0[:, :, x1], rows1[:, :, x0], rows1[:, :, x1]
    top = tl + (tr - tl) * lx
    bottom = bl + (br - bl) * lx
    return top + (bottom - top) * ly


def draw_colored_bounding_boxes(images, boxes, steps):
    """air_model.py:130-158.  images [N,H,W,1], boxes [N,>=3,H,W], steps [N] int -> [N,H,W,3].
    The frame of step s is added to colour channel s (saturating at 1) and subtracted from the other two (saturating
    at 0), for the images that executed more than s steps.  All three channels are carried in one [N,H,W,3] tensor:
    per step, a +1 / -1 / -1 pattern along the channel axis says which saturation applies -- the same two roundings
    per element as the reference's nine tf.where updates, hence bit-identical."""
    rgb = images.expand(-1, -1, -1, 3)
    plus = torch.eye(3, dtype=torch.bool, device=images.device)           # plus[s, c]: channel c gains frame s
    for s in range(3):
        frame = boxes[:, s, :, :, None]                                   # [N,H,W,1], broadcast over the channels
        updated = torch.where(plus[s], (rgb + frame).clamp_max(1.0), (rgb - frame).clamp_min(0.0))
        rgb = torch.where((steps > s).view(-1, 1, 1, 1), updated, rgb)
    return rgb


def visualize_reconstructions(original, reconstruction, st_back, steps, max_steps=3, canvas_size=50, windows_size=28,
                              zoom=2, transformer=None):
    """air_model.py:211-267.  original / reconstruction [N, canvas^2], st_back [N, T, 2, 3] (T <= max_steps executed
    steps, zero-padded like the reference does), steps [N] -> [N, zoom*canvas, 2*zoom*canvas + 4, 3] in [0, 1]."""
    if transformer is None:
        from ..air.transformer import transformer
    N, big = original.shape[0], zoom * canvas_size
    large_original = resize_bilinear(original.reshape(N, canvas_size, canvas_size, 1), (big, big))
    large_reconstruction = resize_bilinear(reconstruction.reshape(N, canvas_size, canvas_size, 1), (big, big))
    st_back = st_back.reshape(N, -1, 2, 3)
    if st_back.shape[1] < max_steps:
        st_back = torch.cat([st_back, st_back.new_zeros(N, max_steps - st_back.shape[1], 2, 3)], dim=1)
    # tf.image.draw_bounding_boxes(zeros, [[0, 0, 1, 1]]): the one-pixel border of the window set to 1
    frame = original.new_zeros(N * max_steps, windows_size, windows_size, 1)
    frame[:, 0, :, :] = 1.0
    frame[:, -1, :, :] = 1.0
    frame[:, :, 0, :] = 1.0
    frame[:, :, -1, :] = 1.0
    boxes = transformer(frame, st_back.reshape(N * max_steps, 6).contiguous(), (big, big))
    boxes = torch.clamp(boxes, 0.0, 1.0).reshape(N, max_steps, big, big)
    boxes = torch.where(boxes > 0.01, torch.ones_like(boxes), torch.zeros_like(boxes))   # sharpen the frames
    steps = steps.to(original.device)
    return torch.cat([draw_colored_bounding_boxes(large_original, boxes, steps),
                      original.new_ones(N, big, 4, 3),
                      draw_colored_bounding_boxes(large_reconstruction, boxes, steps)], dim=2)
