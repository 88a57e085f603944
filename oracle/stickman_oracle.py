"""CPU oracle of the stickman rasteriser — TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

Restates ``make_joint_img`` of the reference (``lib/utils.py:325-513``) in numpy, including the two OpenCV
primitives it calls with the reference's arguments (thickness 1, 8-connected, no anti-aliasing, shift 0):

* ``cv2.line``  — OpenCV 4.x ``drawing.cpp``: ``ThickLine`` with thickness 1 -> ``Line`` -> ``clipLine`` (Cohen-Sutherland
  with the intersection computed in double precision and truncated) -> ``LineIterator`` (left-to-right 8-connected
  Bresenham: ``err = dx - 2*dy``, step along the minor axis when ``err < 0``).
* ``cv2.fillPoly`` — ``CollectPolyEdges`` (outline with ``Line``; edges in 16.16 fixed point whose slope comes from the
  CLIPPED end points and whose start is extrapolated back to the unclipped row; a clip that degenerates to one row
  leaves a vertical edge at the image border) + ``FillEdgeCollection`` (per scan line:
  active edges sorted by x, spans ``[ceil(x_a), floor(x_b)]`` between consecutive pairs).

OpenCV is a third-party dependency of the reference (``environment.yml``: opencv 4.x; this image has 4.13) whose source
is not in ``/root/reference``; the restatement is pinned twice: ``oracle/gen_stickman_golden.py`` runs the UNMODIFIED
reference function (its source text, executed with the real ``cv2``) and this oracle on the same poses and commits the
results (``tests/golden/stickman.npz``), and ``tests/test_stickman.py`` re-checks the primitives against ``cv2`` itself on
tens of thousands of random cases where ``cv2`` is importable.

``fill_poly`` is exact for vertices anywhere (inside, above / below, left / right of the image): 40 000 random
polygons with vertices up to 15-40 pixels outside a 64 x 64 image reproduce ``cv2.fillPoly`` bit for bit.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

XY_SHIFT = 16
XY_ONE = 1 << XY_SHIFT

JointModel = namedtuple("JointModel", "body right_lines left_lines head_lines face rshoulder lshoulder headup")


def get_line_colors(n_lines_per_channel):
    """lib/utils.py:51-66."""
    line_colors = []
    for channel, nr_lines in enumerate(n_lines_per_channel):
        intervall = int(255 // (nr_lines + 1))
        cols = []
        for i in range(nr_lines):
            col = [0, 0, 0]
            col[channel] = (i + 1) * intervall
            cols.append(col)
        line_colors.append(cols)
    return line_colors


def _trunc_div(num: int, den: int) -> int:
    q = abs(num) // abs(den)
    return q if (num >= 0) == (den > 0) else -q


def clip_line(w: int, h: int, x1: int, y1: int, x2: int, y2: int):
    """cv::clipLine on the rectangle [0, w) x [0, h); returns (visible, x1, y1, x2, y2) — the points are modified even
    when the result is invisible, exactly like the in-place C++ routine."""
    right, bottom = w - 1, h - 1

    def code(x, y):
        return (x < 0) + (x > right) * 2 + (y < 0) * 4 + (y > bottom) * 8

    c1, c2 = code(x1, y1), code(x2, y2)
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def draw_line(img: np.ndarray, p0, p1, color: int) -> None:
    """cv2.line(img, p0, p1, color, thickness=1) on a single-channel uint8 image."""
    h, w = img.shape
    ok, x0, y0, x1, y1 = clip_line(w, h, int(p0[0]), int(p0[1]), int(p1[0]), int(p1[1]))
    if not ok:
        return
    dx, dy = x1 - x0, y1 - y0
    if dx < 0:  # LineIterator(left_to_right = true)
        x0, y0, x1, y1 = x1, y1, x0, y0
        dx, dy = -dx, -dy
    sy = 1 if dy >= 0 else -1
    dy = abs(dy)
    steep = dy > dx
    major, minor = (dy, dx) if steep else (dx, dy)
    err, plus, minus = major - 2 * minor, 2 * major, -2 * minor
    x, y = x0, y0
    for _ in range(major + 1):
        img[y, x] = color
        m = err < 0
        err += minus + (plus if m else 0)
        if steep:
            y += sy
            x += 1 if m else 0
        else:
            x += 1
            y += sy if m else 0


def fill_poly(img: np.ndarray, pts, color: int) -> None:
    """cv2.fillPoly(img, [pts], color) for one integer polygon on a single-channel uint8 image."""
    h, w = img.shape
    pts = [(int(x), int(y)) for x, y in pts]
    edges = []
    p0 = (pts[-1][0] << XY_SHIFT, pts[-1][1])
    for px, py in pts:
        p1 = (px << XY_SHIFT, py)
        t0 = ((p0[0] + (XY_ONE >> 1)) >> XY_SHIFT, p0[1])
        t1 = ((p1[0] + (XY_ONE >> 1)) >> XY_SHIFT, p1[1])
        draw_line(img, t0, t1, color)
        p0c, p1c = p0, p1
        if not (0 <= t0[0] < w and 0 <= t1[0] < w and 0 <= t0[1] < h and 0 <= t1[1] < h):
            _, cx0, cy0, cx1, cy1 = clip_line(w, h, t0[0], t0[1], t1[0], t1[1])
            # "use clipped endpoints to create a more accurate PolyEdge": the x's are always replaced, the y's only when
            # the visible part spans more than one row — a degenerate clip therefore leaves a VERTICAL edge at the
            # clipped x (the image border) over the original rows
            if cy0 != cy1:
                p0c, p1c = (cx0 << XY_SHIFT, cy0), (cx1 << XY_SHIFT, cy1)
            else:
                p0c, p1c = (cx0 << XY_SHIFT, p0[1]), (cx1 << XY_SHIFT, p1[1])
        if p0[1] != p1[1]:
            dxe = _trunc_div(p1c[0] - p0c[0], p1c[1] - p0c[1])
            if p0[1] < p1[1]:
                edges.append((p0[1], p1[1], p0c[0] + (p0[1] - p0c[1]) * dxe, dxe))
            else:
                edges.append((p1[1], p0[1], p1c[0] + (p1[1] - p1c[1]) * dxe, dxe))
        p0 = p1
    if not edges:
        return
    y_min = min(e[0] for e in edges)
    y_max = min(max(e[1] for e in edges), h)
    for y in range(max(y_min, 0), y_max):
        xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(xs) - 1, 2):
            x1 = (xs[k] + XY_ONE - 1) >> XY_SHIFT
            x2 = xs[k + 1] >> XY_SHIFT
            if x1 < w and x2 >= 0:
                x1, x2 = max(x1, 0), min(x2, w - 1)
                if x2 >= x1:
                    img[y, x1:x2 + 1] = color


def _ipt(p):
    """tuple(np.int_(p)): C truncation of the float coordinates."""
    return int(np.int_(p[0])), int(np.int_(p[1]))


def make_joint_img(img_shape, joints, joint_model: JointModel, line_colors=None, scale_factor=None) -> np.ndarray:
    """lib/utils.py:325-513 with ``color_channel=None``; ``joints`` float [J, 2] in image coordinates (x, y), negative =
    missing.  Only thickness 1 is restated (every shipped config leaves ``stickman_scale`` unset -> thickness 1)."""
    thickness = int(img_shape[1] // scale_factor) if scale_factor is not None else 1
    if thickness != 1:
        raise NotImplementedError("stickman oracle: thickness 1 only (no shipped config sets stickman_scale)")
    joints = np.asarray(joints)
    imgs = [np.zeros(img_shape[:2], dtype="uint8") for _ in range(3)]

    def valid(idx):
        return bool(np.all(np.greater_equal(joints[list(idx), :], [0.0, 0.0])))

    if len(joint_model.body) > 2:  # 344-354
        body = np.array([joints[p, :] for p in joint_model.body])
        ok = np.all(np.greater_equal(body, [0.0, 0.0]), axis=-1)
        if np.count_nonzero(ok) > 2:
            pts = [_ipt(p) for p in body[ok]]
            for ch, col in enumerate((0, 127, 255)):
                fill_poly(imgs[ch], pts, col)
    for line_nr, line in enumerate(joint_model.right_lines):  # 356-376
        if valid(line):
            a, b = _ipt(joints[line[0]]), _ipt(joints[line[1]])
            if line_colors is not None:
                ch = int(np.nonzero(line_colors[0][line_nr])[0][0])
                draw_line(imgs[ch], a, b, line_colors[0][line_nr][ch])
            else:
                draw_line(imgs[1], a, b, 255)
    for line_nr, line in enumerate(joint_model.left_lines):  # 378-400
        if valid(line):
            a, b = _ipt(joints[line[0]]), _ipt(joints[line[1]])
            if line_colors is not None:
                ch = int(np.nonzero(line_colors[1][line_nr])[0][0])
                draw_line(imgs[ch], a, b, line_colors[1][line_nr][ch])
            else:
                draw_line(imgs[0], a, b, 255)
    if len(joint_model.head_lines) == 0:  # 402-428: throat from the shoulder midpoint to the head
        rs, ls, cn = joints[joint_model.rshoulder], joints[joint_model.lshoulder], joints[joint_model.headup]
        if np.any(np.less(np.stack([rs, ls], axis=-1), [0.0, 0.0])):
            neck = np.asarray([-1.0, -1.0])
        else:
            neck = 0.5 * (rs + ls)
        pts = np.stack([neck, cn], axis=-1).transpose()
        throat_len = 0.0
        if np.all(np.greater_equal(pts, [0.0, 0.0])):
            throat_len = float(np.linalg.norm(pts[0] - pts[1]))
            a, b = _ipt(pts[0]), _ipt(pts[1])
            draw_line(imgs[0], a, b, 127)
            draw_line(imgs[1], a, b, 127)
    else:  # 429-468
        throat_lens = np.zeros(len(joint_model.head_lines))
        for line_nr, line in enumerate(joint_model.head_lines):
            if valid(line):
                throat_lens[line_nr] = np.linalg.norm(joints[line[0]] - joints[line[1]])
                a, b = _ipt(joints[line[0]]), _ipt(joints[line[1]])
                if line_colors is not None:
                    ch = int(np.nonzero(line_colors[2][line_nr])[0][0])
                    draw_line(imgs[ch], a, b, line_colors[2][line_nr][ch])
                else:
                    draw_line(imgs[0], a, b, 127)
                    draw_line(imgs[1], a, b, 127)
        throat_len = float(np.amax(throat_lens)) if throat_lens.size > 0 else 0.0
    for line_nr, line in enumerate(joint_model.face):  # 469-506: face lines shorter than the throat
        if valid(line) and np.linalg.norm(joints[line[0]] - joints[line[1]]) < throat_len:
            a, b = _ipt(joints[line[0]]), _ipt(joints[line[1]])
            if line_colors is not None:
                ch = int(np.nonzero(line_colors[2][line_nr])[0][0])
                draw_line(imgs[ch], a, b, line_colors[2][line_nr][ch])
            else:
                draw_line(imgs[0], a, b, 127)
                draw_line(imgs[1], a, b, 127)
    return np.stack(imgs, axis=-1)
