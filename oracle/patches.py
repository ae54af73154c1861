"""CPU restatement of the 40 x 40 rgb patch observation of the pixel policies - TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

PARITY UNPINNED for the rendering half, and it cannot be pinned here: the reference draws its frame with gym 0.10.9's
pyglet / OpenGL viewer (gym_ballenv/envs/ballenv_env.py:295-314, 357-386 -> gym.envs.classic_control.rendering, not
vendored, not installable, no GL in this image), and its live patch function (examples/ball_cnn_reinforce.py:120-163;
the copy in examples/ball_cnn_ac3.py:248-307 is commented out of its loop, :557) feeds a [H, W, C] tensor to a
torchvision ToPILImage that wants [C, H, W] and slices with Python-2 integer division.  What is restated is the
intended computation:

  frame   500 x 500, white; in draw order (ballenv_env.py:367-376): the agent, a black 30-gon of radius 5
          (rendering.make_circle(res=30), default colour); the goal, the black quad (5,5),(5,-5),(-5,5),(-5,-5) -
          self-intersecting, drawn as GL_QUADS, restated as the triangle fan (v0,v1,v2),(v0,v2,v3); then every obstacle
          in list order (static, then moving), a 30-gon of radius 20, (100,0,0) -> red if obstacle.speed == 0 else
          (0,100,0) -> green (:298-306; GL clamps colours to 1).  A pixel takes the colour of the last geometry whose
          polygon contains its centre; no anti-aliasing.  Row 0 of the array is the top (y = 499).
  patch   the frame padded by width / 2 with white, cropped to width x width around the agent
          (ball_cnn_reinforce.py:130-144), resized to 40 x 40 with PIL's BICUBIC (ball_cnn_reinforce.py:124) or BILINEAR
          (ball_cnn_ac3.py:255) and scaled to [0, 1] float32 [3, 40, 40] (ToTensor).
The resize half IS pinned: resize_u8() restates Pillow's 8-bit two-pass resampling (Resample.c: precompute_coeffs,
normalize_coeffs_8bpc, horizontal pass then vertical pass, 22-bit fixed point) and tests/test_patches.py checks it
against the installed Pillow bit for bit.

Coordinates are rounded to the nearest integer before drawing (the gym ruleset's coordinates are integral)."""
from __future__ import annotations

import math

import numpy as np

FRAME = 500            # _screen_width / _screen_height, ballenv_env.py:11-12
R_OBSTACLE = 20        # radius_rand_person, ballenv_env.py:49
R_AGENT = 5            # radius_ctrl_person, ballenv_env.py:50
GOAL_QUAD = ((5.0, 5.0), (5.0, -5.0), (-5.0, 5.0), (-5.0, -5.0))   # ballenv_env.py:371
WHITE, BLACK, RED, GREEN = 7, 0, 1, 2   # bit 0 = R, bit 1 = G, bit 2 = B at 255


def circle_polygon(radius, res=30):
    """gym.envs.classic_control.rendering.make_circle (gym 0.10.9, restated)."""
    return [(math.cos(2 * math.pi * i / res) * radius, math.sin(2 * math.pi * i / res) * radius) for i in range(res)]


def _in_triangle(px, py, a, b, c):
    d1 = (b[0] - a[0]) * (py - a[1]) - (b[1] - a[1]) * (px - a[0])
    d2 = (c[0] - b[0]) * (py - b[1]) - (c[1] - b[1]) * (px - b[0])
    d3 = (a[0] - c[0]) * (py - c[1]) - (a[1] - c[1]) * (px - c[0])
    return (d1 >= 0 and d2 >= 0 and d3 >= 0) or (d1 <= 0 and d2 <= 0 and d3 <= 0)


def sprite_rows(verts, radius):
    """Row masks of a polygon (triangle fan from its first vertex) centred on an integral point: bit ix of row iy is set
    iff the centre (ix - R + 0.5, iy - R + 0.5) of that pixel lies inside.  iy counts upwards (GL rows)."""
    rows = []
    for iy in range(2 * radius):
        m = 0
        for ix in range(2 * radius):
            px, py = ix - radius + 0.5, iy - radius + 0.5
            if any(_in_triangle(px, py, verts[0], verts[k], verts[k + 1]) for k in range(1, len(verts) - 1)):
                m |= 1 << ix
        rows.append(m)
    return rows


_SPRITES = {}


def sprites():
    if not _SPRITES:
        _SPRITES["agent"] = sprite_rows(circle_polygon(R_AGENT), R_AGENT)
        _SPRITES["goal"] = sprite_rows(GOAL_QUAD, 5)
        _SPRITES["obstacle"] = sprite_rows(circle_polygon(R_OBSTACLE), R_OBSTACLE)
    return _SPRITES


def render_frame(agent, goal, statics, dynamics, dynamic_speeds=None):
    """The frame as colour codes [500, 500] uint8 (row 0 = top)."""
    f = np.full((FRAME, FRAME), WHITE, dtype=np.uint8)
    sp = sprites()

    def draw(rows, radius, ox, oy, colour):
        ox, oy = int(np.rint(ox)), int(np.rint(oy))
        for iy, m in enumerate(rows):
            y = oy - radius + iy
            if not 0 <= y < FRAME:
                continue
            for ix in range(2 * radius):
                c = ox - radius + ix
                if (m >> ix) & 1 and 0 <= c < FRAME:
                    f[FRAME - 1 - y, c] = colour

    draw(sp["agent"], R_AGENT, agent[0], agent[1], BLACK)
    draw(sp["goal"], 5, goal[0], goal[1], BLACK)
    for o in statics:
        draw(sp["obstacle"], R_OBSTACLE, o[0], o[1], RED)
    for j, o in enumerate(dynamics):
        speed = 1 if dynamic_speeds is None else dynamic_speeds[j]
        draw(sp["obstacle"], R_OBSTACLE, o[0], o[1], RED if speed == 0 else GREEN)
    return f


def codes_to_rgb(codes):
    return np.stack([np.where(codes & 1, 255, 0), np.where(codes & 2, 255, 0), np.where(codes & 4, 255, 0)],
                    axis=-1).astype(np.uint8)


def crop(frame_codes, agent, width=100):
    """ball_cnn_reinforce.py:130-144: pad by width / 2 with white, take [agent_y - span, agent_y + span) x
    [agent_x - span, agent_x + span) of the padded frame (agent_y counted from the top)."""
    span = width // 2
    ax, ay = int(np.rint(agent[0])), int(np.rint(agent[1]))
    padded = np.full((FRAME + 2 * span, FRAME + 2 * span), WHITE, dtype=np.uint8)
    padded[span:span + FRAME, span:span + FRAME] = frame_codes
    px, py = ax + span, FRAME - ay + span
    out = np.full((width, width), WHITE, dtype=np.uint8)
    r0, r1, c0, c1 = py - span, py + span, px - span, px + span
    rr0, rr1, cc0, cc1 = max(r0, 0), min(r1, padded.shape[0]), max(c0, 0), min(c1, padded.shape[1])
    if rr1 > rr0 and cc1 > cc0:
        out[rr0 - r0:rr1 - r0, cc0 - c0:cc1 - c0] = padded[rr0:rr1, cc0:cc1]
    return out


# ---- Pillow's 8-bit resampling, restated (src/libImaging/Resample.c) ------------------------------------------------
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def coefficients(in_size, out_size, interp):
    """precompute_coeffs + normalize_coeffs_8bpc: (ksize, bounds [out][2] = (first, count), kk [out][ksize] int32)."""
    filt, fsupport = FILTERS[interp]
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w[i] for i in range(xmax)) if xmax else 0.0
        # (Pillow accumulates in this order: a running double sum)
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(img, bounds, kk, axis):
    """One resampling pass over `axis` of a uint8 [H, W, C] image."""
    img = img.astype(np.int64)
    n_out = bounds.shape[0]
    shape = list(img.shape)
    shape[axis] = n_out
    out = np.zeros(shape, dtype=np.uint8)
    for xx in range(n_out):
        xmin, cnt = int(bounds[xx, 0]), int(bounds[xx, 1])
        sl = [slice(None)] * 3
        sl[axis] = slice(xmin, xmin + cnt)
        k = kk[xx, :cnt].astype(np.int64)
        kshape = [1, 1, 1]
        kshape[axis] = cnt
        ss = (1 << (PRECISION_BITS - 1)) + (img[tuple(sl)] * k.reshape(kshape)).sum(axis=axis)
        # the accumulators are 32-bit ints in Pillow: wrap like they do (never happens for normalised kernels)
        ss = ((ss + (1 << 31)) % (1 << 32)) - (1 << 31)
        v = np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)
        osl = [slice(None)] * 3
        osl[axis] = xx
        out[tuple(osl)] = v
    return out


def resize_u8(img, out_size, interp):
    """ImagingResample of a uint8 [H, W, 3] image to [out, out, 3]: horizontal pass, then vertical pass."""
    _, bh, kh = coefficients(img.shape[1], out_size, interp)
    tmp = _pass(img, bh, kh, axis=1)
    _, bv, kv = coefficients(img.shape[0], out_size, interp)
    return _pass(tmp, bv, kv, axis=0)


def extract_patch_u8(agent, goal, statics, dynamics, width=100, out_size=40, interp="bicubic", dynamic_speeds=None):
    """uint8 [3, out, out] patch of one environment."""
    codes = crop(render_frame(agent, goal, statics, dynamics, dynamic_speeds), agent, width)
    return np.ascontiguousarray(resize_u8(codes_to_rgb(codes), out_size, interp).transpose(2, 0, 1))


def extract_patch(agent, goal, statics, dynamics, width=100, out_size=40, interp="bicubic", dynamic_speeds=None):
    """float32 [3, out, out] in [0, 1] (ToTensor: uint8 / 255 in float32)."""
    return extract_patch_u8(agent, goal, statics, dynamics, width, out_size, interp, dynamic_speeds).astype(np.float32) / np.float32(255)
