"""fast_artistic_video/vr_helper.lua -- perspective border warp maps and the cube -> equirectangular map (host side,
computed once per stream).  Double arithmetic in the reference's expression order; Lua's float-valued `for` loops and
the float -> index truncation are reproduced.  Maps are [2,H,W] with channel 0 = dy, 1 = dx and sentinel 99999 where
no neighbour pixel maps (vr_helper.lua:10,32,55,79), which the warp turns into exact zeros."""
from __future__ import annotations

import math

import numpy as np


def _lua_range(start: float, stop: float):
    """Values of a Lua numeric for loop `for x=start,stop do` (step 1, repeated addition)."""
    out, x = [], float(start)
    while x <= stop:
        out.append(x)
        x = x + 1.0
    return out


def _width(height, oversize_h):
    width = height / 2 / ((2 * oversize_h + height) / height)
    max_resize_factor = (width + oversize_h) / width
    return width - (max_resize_factor - 1) / max_resize_factor * oversize_h


def make_perspective_warp_map_left(height, crop_w, orig_width, oversize_h=None, oversize_w=None):  # vr_helper.lua:3-24
    oversize_h = crop_w / 2 if oversize_h is None else oversize_h
    oversize_w = crop_w / 2 if oversize_w is None else oversize_w
    width = _width(height, oversize_h)
    m = np.full((2, height, orig_width), 99999.0)
    mid_y = height / 2
    y = np.arange(1, height + 1, dtype=np.float64)
    for x in _lua_range(width - crop_w + 1, width):
        rh, rw = (x + oversize_h) / width, (x + oversize_w) / width
        col = int(x - (width - crop_w) + orig_width - crop_w) - 1
        m[0, :, col] = (mid_y - y) * (-1 / rh + 1)
        m[1, :, col] = (width - x - oversize_w) * (rw - 1) / rw - orig_width + crop_w
    return m


def make_perspective_warp_map_right(height, crop_w, org_width, oversize_h=None, oversize_w=None):  # :26-47
    oversize_h = crop_w / 2 if oversize_h is None else oversize_h
    oversize_w = crop_w / 2 if oversize_w is None else oversize_w
    width = _width(height, oversize_h)
    m = np.full((2, height, org_width), 99999.0)
    mid_y = height / 2
    y = np.arange(1, height + 1, dtype=np.float64)
    for x in range(1, crop_w + 1):
        rh, rw = (width - x + oversize_h) / width, (width - x + oversize_w) / width
        m[0, :, x - 1] = (mid_y - y) * (-1 / rh + 1)
        m[1, :, x - 1] = -(x - oversize_w) * (rw - 1) / rw + org_width - crop_w
    return m


def make_perspective_warp_map_top(width, crop_h, orig_height, oversize_w=None, oversize_h=None):  # :49-71
    oversize_h = crop_h / 2 if oversize_h is None else oversize_h
    oversize_w = crop_h / 2 if oversize_w is None else oversize_w
    height = _width(width, oversize_w)
    m = np.full((2, orig_height, width), 99999.0)
    mid_x = width / 2
    x = np.arange(1, width + 1, dtype=np.float64)
    for yv in _lua_range(height - crop_h + 1, height):
        rw, rh = (yv + oversize_w) / height, (yv + oversize_h) / height
        row = int(yv - (height - crop_h) + orig_height - crop_h) - 1
        m[0, row, :] = (height - yv - oversize_h) * (rh - 1) / rh - orig_height + crop_h
        m[1, row, :] = (mid_x - x) * (-1 / rw + 1)
    return m


def make_perspective_warp_map_bottom(width, crop_h, orig_height, oversize_w=None, oversize_h=None):  # :74-92
    oversize_h = crop_h / 2 if oversize_h is None else oversize_h
    oversize_w = crop_h / 2 if oversize_w is None else oversize_w
    height = _width(width, oversize_w)
    m = np.full((2, orig_height, width), 99999.0)
    mid_x = width / 2
    x = np.arange(1, width + 1, dtype=np.float64)
    for yv in range(1, crop_h + 1):
        rw, rh = (height - yv + oversize_w) / height, (height - yv + oversize_h) / height
        m[0, yv - 1, :] = -(yv - oversize_h) * (rh - 1) / rh + orig_height - crop_h
        m[1, yv - 1, :] = (mid_x - x) * (-1 / rw + 1)
    return m


def make_cube_to_equirectangular_map(w_plus_overlap, h_plus_overlap, overlap_w, overlap_h, out_w, out_h):  # :95-184
    cw, ch = w_plus_overlap - overlap_w, h_plus_overlap - overlap_h
    j = np.arange(out_h, dtype=np.float64)[:, None]
    i = np.arange(out_w, dtype=np.float64)[None, :]
    theta = (1 - (j / out_h)) * math.pi
    phi = (i / out_w) * 2 * math.pi
    x = np.sin(phi) * np.sin(theta) * -1
    y = np.cos(theta) * np.ones_like(phi)
    z = np.cos(phi) * np.sin(theta) * -1
    a = np.maximum(np.maximum(np.abs(x), np.abs(y)), np.abs(z))
    xa, ya, za = x / a, y / a, z / a
    conds = [xa == 1, xa == -1, ya == 1, ya == -1, za == 1, za == -1]  # first match wins, as the elseif chain
    xpix = np.select(conds, [((za + 1) / 2 - 1) * cw, ((za + 1) / 2) * cw, ((xa + 1) / 2) * cw, ((xa + 1) / 2) * cw,
                             ((xa + 1) / 2) * cw, ((xa + 1) / 2 - 1) * cw], 0.0)
    xoff = np.select(conds, [2.0 * w_plus_overlap, 1.0 * w_plus_overlap, 5.0 * w_plus_overlap, 4.0 * w_plus_overlap,
                             0.0 * w_plus_overlap, 3.0 * w_plus_overlap], 0.0)
    ypix = np.select(conds, [((ya + 1) / 2) * ch, ((ya + 1) / 2) * ch, ((za + 1) / 2 - 1) * ch, ((za + 1) / 2) * ch,
                             ((ya + 1) / 2) * ch, ((ya + 1) / 2) * ch], 0.0)
    xpix = np.abs(xpix) + xoff + overlap_w / 2
    ypix = np.abs(ypix) + 0 + overlap_h / 2
    return np.stack([ypix - j, xpix - i])


# fast_artistic_video/utils.lua:179-213 (values only; the reference expands them to c x h x w views)
def make_gradient_mask_h_inc(h):
    return np.arange(1, h + 1, dtype=np.float64) / (h + 1)


def make_gradient_mask_h_dec(h):
    return np.arange(h, 0, -1, dtype=np.float64) / (h + 1)


def make_gradient_mask_w_inc(w):
    return (np.arange(1, w + 1, dtype=np.float64) / (w + 1)).astype(np.float32).astype(np.float64)  # :float() in :201


def make_gradient_mask_w_dec(w):
    return np.arange(w, 0, -1, dtype=np.float64) / (w + 1)
