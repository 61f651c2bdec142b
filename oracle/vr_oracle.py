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


# ---------------------------------------------------------------------------------------------------------------------
# The cube-map DRIVER (fast_artistic_video_vr.lua:154-302, 454-559) + the frame loop that calls it
# (fast_artistic_video_core.lua:161-229), restated statement by statement with numpy.  Tensor types are the reference's:
# float32 wherever the Lua code holds a `dtype` (CudaTensor) value, float64 for the DoubleTensor gradient masks until the
# line that casts them.  The stylization network itself is injected (`net.run_image(content)` /
# `net.run_next_image_prior(content, prior, cert, flow_mask)`), so the driver logic is checked independently of it.
# ---------------------------------------------------------------------------------------------------------------------
_f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
PROC_ORDER = [6, 1, 2, 5, 3, 4]  # :103


def _warp(img, m):  # utils.warp_image(img, map, 'torch.CudaTensor')  (utils.lua:141-144)
    return pyoracle.warp_bdhw(_f32(img), _f32(m))


class VRRef:
    """State of fast_artistic_video_vr.lua:76-94 + its callbacks.  Faces are 3 x hplus x wplus float32 arrays."""

    def __init__(self, hplus, wplus, overlap_w, overlap_h, median_filter=3, out_equi_w=768, out_equi_h=768,
                 smooth_certainty=False):
        self.hplus, self.wplus, self.ow, self.oh, self.mf = hplus, wplus, overlap_w, overlap_h, median_filter
        self.smooth_certainty = smooth_certainty
        ones = np.ones((1, hplus, wplus), np.float32)
        # :170-179
        self.map_left = _f32(perspective_left(hplus, overlap_w, wplus)); self.mask_left = _warp(ones, self.map_left)
        self.map_top = _f32(perspective_top(wplus, overlap_h, hplus)); self.mask_top = _warp(ones, self.map_top)
        self.map_bottom = _f32(perspective_bottom(wplus, overlap_h, hplus)); self.mask_bottom = _warp(ones, self.map_bottom)
        self.map_right = _f32(perspective_right(hplus, overlap_w, wplus)); self.mask_right = _warp(ones, self.map_right)
        msum = ((self.mask_left + self.mask_right).astype(np.float32) + self.mask_top).astype(np.float32) + self.mask_bottom
        msum = msum.astype(np.float32)
        self.mask_all_div = np.maximum(msum, np.float32(1))  # torch.cmax(.., 1)
        self.mask_all = np.minimum(msum, np.float32(1))      # torch.cmin(.., 1)
        gh, gw = overlap_h - 10, overlap_w - 10              # :181-182
        # utils.make_gradient_mask_* (utils.lua:179-213): i/(n+1) in double; w_inc is cast :float() (:202) and back :double() (:185)
        w_dec = np.arange(gw, 0, -1, dtype=np.float64) / (gw + 1)
        w_inc = (np.arange(1, gw + 1, dtype=np.float64) / (gw + 1)).astype(np.float32).astype(np.float64)
        h_dec = np.arange(gh, 0, -1, dtype=np.float64) / (gh + 1)
        h_inc = np.arange(1, gh + 1, dtype=np.float64) / (gh + 1)
        Z = np.zeros
        self.g_left = np.concatenate([np.broadcast_to(w_dec[None, None, :], (1, hplus, gw)), Z((1, hplus, wplus - gw))], 2)     # :184
        self.g_right = np.concatenate([Z((1, hplus, wplus - gw)), np.broadcast_to(w_inc[None, None, :], (1, hplus, gw))], 2)    # :185
        self.g_top = np.concatenate([np.broadcast_to(h_dec[None, :, None], (1, gh, wplus)), Z((1, hplus - gh, wplus))], 1)      # :186
        self.g_bottom = np.concatenate([Z((1, hplus - gh, wplus)), np.broadcast_to(h_inc[None, :, None], (1, gh, wplus))], 1)   # :187
        self.g_all = np.maximum(np.maximum(self.g_left, self.g_right), np.maximum(self.g_top, self.g_bottom))                   # :188
        self.g_left_right = np.maximum(self.g_left, self.g_right)                                                               # :189
        r = median_filter // 2
        self.equi_map = _f32(cube_to_equirect(hplus - 2 * r, wplus - 2 * r, overlap_w - r, overlap_h - r, out_equi_w, out_equi_h))  # :193-194
        self.last_segments, self.prev_last_segments = {}, {}

    # :204-237 (cert_frame: the loaded occlusion PGM / 255, or None for i < 7)
    def load_cert(self, mode, cert_frame):
        cb = np.zeros((1, self.hplus, self.wplus), np.float32)
        if mode in (1, 3, 4, 5):
            cb = np.maximum(cb, self.mask_left)
        if mode in (2, 3, 4, 5):
            cb = np.maximum(cb, self.mask_right)
        if mode in (4, 5):
            cb = np.maximum(cb, self.mask_top)
            cb = np.maximum(cb, self.mask_bottom)
        return np.maximum(_f32(cert_frame), cb) if cert_frame is not None else cb

    # :239-302 (flow: the face's backward flow in (dy,dx) layout, or None for i < 7; cert: min-filtered certainty 1xHxW)
    def make_last_frame_warped(self, mode, flow, cert):
        ls, div = self.last_segments, self.mask_all_div
        border = np.zeros((3, self.hplus, self.wplus), np.float32)
        grad = None
        add = lambda a, b: (a + b).astype(np.float32)
        cdiv = lambda a: (a / div).astype(np.float32)
        if mode == 1:
            border, grad = _warp(ls[1], self.map_left), self.g_right
        elif mode == 2:
            border, grad = _warp(ls[1], self.map_right), self.g_left
        elif mode == 3:
            border = add(_warp(ls[2], self.map_left), _warp(ls[3], self.map_right)); grad = self.g_left_right
        elif mode == 4:
            border = cdiv(_warp(rot(ls[2], 1), self.map_left))
            border = add(border, cdiv(_warp(rot(ls[3], 2), self.map_right)))
            border = add(border, cdiv(_warp(ls[4], self.map_top)))
            border = add(border, cdiv(_warp(rot(ls[1], 3), self.map_bottom)))
            grad = self.g_all
        elif mode == 5:
            border = cdiv(_warp(rot(ls[2], 2), self.map_left))
            border = add(border, cdiv(_warp(rot(ls[3], 1), self.map_right)))
            border = add(border, cdiv(_warp(rot(ls[1], 3), self.map_top)))
            border = add(border, cdiv(_warp(ls[4], self.map_bottom)))
            grad = self.g_all
        if flow is not None:  # i >= 7  (:275-291)
            lfw = _warp(self.prev_last_segments[mode + 1], flow)
            cert_inv = (np.float32(1) - _f32(cert).reshape(1, self.hplus, self.wplus)).astype(np.float32)
            if mode == 0:
                result = lfw
            else:
                gm = [self.g_right, self.g_left, self.g_left_right, self.g_all, self.g_all][mode - 1].astype(np.float32)
                masks = [self.mask_left, self.mask_right, add(self.mask_left, self.mask_right), self.mask_all, self.mask_all][mode - 1]
                mask = (np.maximum(gm, (np.ceil(gm) * cert_inv).astype(np.float32)) * masks).astype(np.float32)  # :288
                anti = (np.float32(1) - mask).astype(np.float32)                                                  # :289
                result = add((lfw * anti).astype(np.float32), (border * mask).astype(np.float32))                 # :290
        else:
            result = border
        if self.smooth_certainty:  # :296-297 (gradMask is nil for mode 0: the reference raises there)
            if grad is None:
                raise RuntimeError("attempt to index a nil value (gradMask), fast_artistic_video_vr.lua:297")
            fm = np.maximum(np.sign(np.maximum((grad.astype(np.float32) - np.float32(0.5)).astype(np.float32), np.float32(0))), np.float32(0.25))
            return result, fm.astype(np.float32)
        return result, None

    # :454-509
    def blend_other_sides(self):
        ls = self.last_segments
        anti = (1.0 - self.g_all).astype(np.float32)[0]  # csub in DOUBLE, then :type(dtype)  (:456)
        mask = self.g_all.astype(np.float32)[0]           # :457
        plan = {1: [(2, self.map_right, 0), (3, self.map_left, 0), (5, self.map_bottom, 3), (6, self.map_top, 3)],
                2: [(1, self.map_left, 0), (4, self.map_right, 0), (5, self.map_bottom, 2), (6, self.map_top, 1)],
                3: [(1, self.map_right, 0), (4, self.map_left, 0), (5, self.map_bottom, 1), (6, self.map_top, 2)],
                4: [(2, self.map_left, 0), (3, self.map_right, 0), (5, self.map_bottom, 0), (6, self.map_top, 0)],
                5: [(1, self.map_bottom, 3), (2, self.map_left, 1), (3, self.map_right, 2), (4, self.map_top, 0)],
                6: [(1, self.map_top, 3), (2, self.map_left, 2), (3, self.map_right, 1), (4, self.map_bottom, 0)]}
        out = {}
        for face, sides in plan.items():
            acc = None
            for s, m, r in sides:  # combineSides :146-152
                q = (_warp(rot(ls[s], r), m) / self.mask_all_div).astype(np.float32)
                acc = q if acc is None else (acc + q).astype(np.float32)
            out[face] = ((ls[face] * anti[None]).astype(np.float32) + (acc * mask[None]).astype(np.float32)).astype(np.float32)
        return out

    # :511-559 (returns dict(equi=..., cubemap=...) after the sixth face, else None)
    def save_image(self, mode, frame):
        self.last_segments[mode + 1] = _f32(frame)
        if mode != 5:
            return None
        self.prev_last_segments = self.blend_other_sides()
        sides = {j: (median_filter(self.prev_last_segments[j], self.mf) if self.mf > 0 else self.prev_last_segments[j])
                 for j in range(1, 7)}
        ow = self.ow // 2 - self.mf // 2  # oversize_w (:515); overlap assumed even as the Lua slices need integers
        oh = self.oh // 2 - self.mf // 2
        strip = np.concatenate([sides[1], sides[2], sides[3], sides[4], rot(sides[5], 3), rot(sides[6], 3)], 2)
        equi = _warp(strip, self.equi_map)  # :543
        # :548-553: {oversize+1, hplus-oversize} are 1-based inclusive indices into the MEDIAN-FILTERED face
        crop = lambda t: t[:, oh:self.hplus - oh, ow:self.wplus - ow]
        cubemap = np.concatenate([crop(sides[4]), crop(sides[1]), rot(crop(sides[5]), 1), rot(crop(sides[6]), 2),
                                  crop(sides[3]), crop(sides[2])], 2)
        return dict(equi=equi, cubemap=cubemap)


def run_vr_clip(ref: VRRef, net, frames, flows, certs, min_filter_r=7, create_inconsistent=False):
    """run_fast_neural_video (fast_artistic_video_core.lua:189-229) with the VR callbacks.
    frames[f][face] (face = 1..6 file index) 3xSxS [0,1]; flows[f][face] backward flow (dy,dx) of VR frame f >= 1 (0-based);
    certs[f][face] occlusion PGM / 255.  Returns the per-VR-frame outputs of save_image."""
    outs = []
    n = len(frames)
    for i in range(1, 6 * n + 1):
        mode, f = (i - 1) % 6, (i - 1) // 6
        face = PROC_ORDER[mode]
        img = _f32(frames[f][face])
        temporal = i >= 7 and not create_inconsistent  # :227, :275
        if (i % 6 == 1) if create_inconsistent else (i == 1):  # func_is_single_image :304-310
            styl = net.run_image(img)
        else:
            cert = ref.load_cert(mode, certs[f][face] if temporal else None)
            cert = pyoracle.min_filter(cert[0], min_filter_r)[None]  # core.lua:207
            prior, fmask = ref.make_last_frame_warped(mode, _f32(flows[f][face]) if temporal else None, cert)
            styl = net.run_next_image_prior(img, prior, cert[0], None if fmask is None else fmask[0])
        res = ref.save_image(mode, _f32(styl))
        if res is not None:
            outs.append(res)
    return outs
