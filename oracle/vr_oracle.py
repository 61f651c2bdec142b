"""Literal (loop-for-loop) Python restatement of fast_artistic_video/vr_helper.lua and of the VR post-processing
(utils.median_filter, combineSides / blend_other_sides).  TEST INFRASTRUCTURE ONLY.  Lua semantics reproduced:
1-based indices, double arithmetic, float-valued numeric `for`, double -> index truncation.  PARITY UNPINNED (no Lua)."""
from __future__ import annotations

import math

import numpy as np

from . import pyoracle


def _lua_for(start, stop):
    x = float(start)
    while x <= stop:
        yield x
        x = x + 1.0


def perspective_left(height, crop_w, orig_width):  # vr_helper.lua:3-24
    oversize_h = crop_w / 2; oversize_w = crop_w / 2
    width = height / 2 / ((2 * oversize_h + height) / height)
    max_resize_factor = (width + oversize_h) / width
    width = width - (max_resize_factor - 1) / max_resize_factor * oversize_h
    m = np.full((2, height, orig_width), 99999.0)
    mid_y = height / 2
    for x in _lua_for(width - crop_w + 1, width):
        resize_factor_h = (x + oversize_h) / width
        resize_factor_w = (x + oversize_w) / width
        for y in range(1, height + 1):
            col = int(x - (width - crop_w) + orig_width - crop_w)
            m[0][y - 1][col - 1] = (mid_y - y) * (-1 / resize_factor_h + 1)
            m[1][y - 1][col - 1] = (width - x - oversize_w) * (resize_factor_w - 1) / resize_factor_w - orig_width + crop_w
    return m


def perspective_right(height, crop_w, org_width):  # :26-47
    oversize_h = crop_w / 2; oversize_w = crop_w / 2
    width = height / 2 / ((2 * oversize_h + height) / height)
    max_resize_factor = (width + oversize_h) / width
    width = width - (max_resize_factor - 1) / max_resize_factor * oversize_h
    m = np.full((2, height, org_width), 99999.0)
    mid_y = height / 2
    for x in range(1, crop_w + 1):
        resize_factor_h = (width - x + oversize_h) / width
        resize_factor_w = (width - x + oversize_w) / width
        for y in range(1, height + 1):
            m[0][y - 1][x - 1] = (mid_y - y) * (-1 / resize_factor_h + 1)
            m[1][y - 1][x - 1] = -(x - oversize_w) * (resize_factor_w - 1) / resize_factor_w + org_width - crop_w
    return m


def perspective_top(width, crop_h, orig_height):  # :49-71
    oversize_h = crop_h / 2; oversize_w = crop_h / 2
    height = width / 2 / ((2 * oversize_w + width) / width)
    max_resize_factor = (height + oversize_w) / height
    height = height - (max_resize_factor - 1) / max_resize_factor * oversize_w
    m = np.full((2, orig_height, width), 99999.0)
    mid_x = width / 2
    for y in _lua_for(height - crop_h + 1, height):
        resize_factor_w = (y + oversize_w) / height
        resize_factor_h = (y + oversize_h) / height
        for x in range(1, width + 1):
            row = int(y - (height - crop_h) + orig_height - crop_h)
            m[0][row - 1][x - 1] = (height - y - oversize_h) * (resize_factor_h - 1) / resize_factor_h - orig_height + crop_h
            m[1][row - 1][x - 1] = (mid_x - x) * (-1 / resize_factor_w + 1)
    return m


def perspective_bottom(width, crop_h, orig_height):  # :74-92
    oversize_h = crop_h / 2; oversize_w = crop_h / 2
    height = width / 2 / ((2 * oversize_w + width) / width)
    max_resize_factor = (height + oversize_w) / height
    height = height - (max_resize_factor - 1) / max_resize_factor * oversize_w
    m = np.full((2, orig_height, width), 99999.0)
    mid_x = width / 2
    for y in range(1, crop_h + 1):
        resize_factor_w = (height - y + oversize_w) / height
        resize_factor_h = (height - y + oversize_h) / height
        for x in range(1, width + 1):
            m[0][y - 1][x - 1] = -(y - oversize_h) * (resize_factor_h - 1) / resize_factor_h + orig_height - crop_h
            m[1][y - 1][x - 1] = (mid_x - x) * (-1 / resize_factor_w + 1)
    return m


def cube_to_equirect(w_plus_overlap, h_plus_overlap, overlap_w, overlap_h, out_w, out_h):  # :95-184
    m = np.zeros((2, out_h, out_w))
    cw, ch = w_plus_overlap - overlap_w, h_plus_overlap - overlap_h
    for j in range(out_h):
        v = 1 - (j / out_h)
        theta = v * math.pi
        for i in range(out_w):
            u = i / out_w
            phi = u * 2 * math.pi
            x = math.sin(phi) * math.sin(theta) * -1
            y = math.cos(theta)
            z = math.cos(phi) * math.sin(theta) * -1
            a = max(abs(x), abs(y), abs(z))
            xa, ya, za = x / a, y / a, z / a
            if xa == 1:
                xp, xo, yp = ((za + 1) / 2 - 1) * cw, 2 * w_plus_overlap, ((ya + 1) / 2) * ch
            elif xa == -1:
                xp, xo, yp = ((za + 1) / 2) * cw, 1 * w_plus_overlap, ((ya + 1) / 2) * ch
            elif ya == 1:
                xp, xo, yp = ((xa + 1) / 2) * cw, 5 * w_plus_overlap, ((za + 1) / 2 - 1) * ch
            elif ya == -1:
                xp, xo, yp = ((xa + 1) / 2) * cw, 4 * w_plus_overlap, ((za + 1) / 2) * ch
            elif za == 1:
                xp, xo, yp = ((xa + 1) / 2) * cw, 0 * w_plus_overlap, ((ya + 1) / 2) * ch
            elif za == -1:
                xp, xo, yp = ((xa + 1) / 2 - 1) * cw, 3 * w_plus_overlap, ((ya + 1) / 2) * ch
            else:
                xp, xo, yp = 0, 0, 0
            xp = abs(xp) + xo + overlap_w / 2
            yp = abs(yp) + 0 + overlap_h / 2
            m[0][j][i] = yp - j
            m[1][j][i] = xp - i
    return m


def median_filter(img: np.ndarray, r: int) -> np.ndarray:
    """utils.median_filter (utils.lua:151-159): unfold r x r, torch median (lower median) over the window."""
    C, H, W = img.shape
    out = np.empty((C, H - r + 1, W - r + 1), np.float32)
    k = (r * r - 1) // 2
    for y in range(H - r + 1):
        for x in range(W - r + 1):
            win = img[:, y:y + r, x:x + r].reshape(C, -1)
            out[:, y, x] = np.sort(win, axis=1)[:, k]
    return out


def rot(t, code):  # fast_artistic_video_vr.lua:134-144
    if code == 0:
        return t
    if code == 1:
        return np.ascontiguousarray(np.transpose(t, (0, 2, 1))[:, ::-1, :])
    if code == 2:
        return np.ascontiguousarray(np.transpose(t, (0, 2, 1))[:, :, ::-1])
    return np.ascontiguousarray(t[:, ::-1, ::-1])


def blend_sides(base, sides, maps, rots, div, mask):
    """combineSides (:146-152) + result = base*anti_mask + borders*mask (:456-466), fp32 in reference order."""
    acc = None
    for s, m, r in zip(sides, maps, rots):
        q = (pyoracle.warp_bdhw(rot(s, r), m) / div[None]).astype(np.float32)
        acc = q if acc is None else (acc + q).astype(np.float32)
    am = (np.float32(1.0) - mask).astype(np.float32)
    return ((base * am[None]).astype(np.float32) + (acc * mask[None]).astype(np.float32)).astype(np.float32)
