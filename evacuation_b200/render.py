"""Rendering of one environment's state / recorded trajectory (src/env/env/env.py:173-324: `render` -> PNG,
`save_animation` -> GIF), host-side and off the device path.

The reference draws with matplotlib and writes its GIFs through matplotlib's Pillow writer.  matplotlib is not a dependency
of this package (and is absent from the B200 image), so the same scene is rasterised directly with Pillow: the arena
(dashed grey boundary of [-w, w] x [-h, h] inside a [-1.1 w, 1.1 w] x [-1.1 h, 1.1 h] view), the exiting zone (half disc of
radius `to_exit` above the exit, green, alpha 0.2), the escaping zone (half disc of radius `to_escape`, white), the exit
marker (green X), the following zone (disc of radius `to_leader` around the agent, blue, alpha 0.1), the pedestrians as dots
coloured by status in matplotlib's default colour cycle in `Status.all()` order (VISCEK tab:blue, FOLLOWER tab:orange,
EXITING tab:green, ESCAPED tab:red), the agent as a red +, and the title line.  500 x 500 pixels = the reference's
`figsize=(5, 5)` at 100 dpi; GIF frames last 20 ms like `FuncAnimation(interval=20)`.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np

SIZE = 500
_MARGIN_TOP = 36  # room for the title line(s)
# Status.all() order with matplotlib's default colour cycle (env.py:207-216)
STATUS_COLORS = {1: (31, 119, 180), 2: (255, 127, 14), 3: (44, 160, 44), 4: (214, 39, 40)}
_GREEN, _BLUE, _RED, _GREY = (0, 128, 0), (0, 0, 255), (255, 0, 0), (128, 128, 128)


def _pil():
    try:
        from PIL import Image, ImageDraw
    except ImportError as exc:  # pragma: no cover
        raise ImportError("render / save_animation need Pillow (the reference needs it too: its GIFs are written by matplotlib's "
                          "Pillow writer); the recorded trajectory is available as env.unwrapped.pedestrians.memory / agent.memory") from exc
    return Image, ImageDraw


class _View:
    """World -> pixel mapping of the reference's axes: x in [-1.1 w, 1.1 w], y in [-1.1 h, 1.1 h], y up."""

    def __init__(self, width: float, height: float, size: int = SIZE):
        self.w, self.h, self.size = float(width), float(height), size
        self.plot = size - _MARGIN_TOP - 8  # square plot area below the title
        self.x0, self.y0 = (size - self.plot) // 2, _MARGIN_TOP
        self.sx, self.sy = self.plot / (2.2 * self.w), self.plot / (2.2 * self.h)

    def px(self, x, y):
        return (self.x0 + (np.asarray(x, dtype=np.float64) + 1.1 * self.w) * self.sx,
                self.y0 + (1.1 * self.h - np.asarray(y, dtype=np.float64)) * self.sy)

    def box(self, cx, cy, r):
        (x, y) = self.px(cx, cy)
        return [float(x - r * self.sx), float(y - r * self.sy), float(x + r * self.sx), float(y + r * self.sy)]


def _dashed(draw, p0, p1, fill, dash=6, gap=4):
    (x0, y0), (x1, y1) = p0, p1
    n = max(1, int(np.hypot(x1 - x0, y1 - y0) // (dash + gap)))
    for i in range(n + 1):
        a = i * (dash + gap) / max(np.hypot(x1 - x0, y1 - y0), 1e-9)
        b = min(1.0, a + dash / max(np.hypot(x1 - x0, y1 - y0), 1e-9))
        if a >= 1.0:
            break
        draw.line([(x0 + (x1 - x0) * a, y0 + (y1 - y0) * a), (x0 + (x1 - x0) * b, y0 + (y1 - y0) * b)], fill=fill, width=1)


def draw_frame(positions, statuses, agent_position, *, width=1.0, height=1.0, exit_position=(0.0, -1.0), to_exit=0.4, to_escape=0.01,
               to_leader=0.2, title: str = ""):
    """One frame as a PIL RGB image.  positions [N,2] (any float dtype), statuses [N] with the reference's enum values 1..4."""
    Image, ImageDraw = _pil()
    view = _View(width, height)
    img = Image.new("RGBA", (SIZE, SIZE), (255, 255, 255, 255))
    overlay = Image.new("RGBA", (SIZE, SIZE), (0, 0, 0, 0))
    od = ImageDraw.Draw(overlay)
    ex, ey = float(exit_position[0]), float(exit_position[1])
    ax_, ay_ = float(agent_position[0]), float(agent_position[1])
    # exiting zone: upper half disc (Wedge 0..180 degrees), green alpha 0.2; PIL angles run clockwise from 3 o'clock
    od.pieslice(view.box(ex, ey, to_exit), 180, 360, fill=_GREEN + (51,))
    # following zone around the agent, blue alpha 0.1
    od.ellipse(view.box(ax_, ay_, to_leader), fill=_BLUE + (26,))
    img = Image.alpha_composite(img, overlay)
    d = ImageDraw.Draw(img)
    d.pieslice(view.box(ex, ey, max(to_escape, 1.5 / view.sx)), 180, 360, fill=(255, 255, 255, 255))  # escaping zone (>= 1.5 px so it shows)
    # arena boundary, dashed grey
    corners = [view.px(-width, -height), view.px(width, -height), view.px(width, height), view.px(-width, height)]
    corners = [(float(x), float(y)) for x, y in corners]
    for i in range(4):
        _dashed(d, corners[i], corners[(i + 1) % 4], _GREY + (255,))
    # exit marker: green X
    cx, cy = (float(v) for v in view.px(ex, ey))
    d.line([(cx - 5, cy - 5), (cx + 5, cy + 5)], fill=_GREEN + (255,), width=3)
    d.line([(cx - 5, cy + 5), (cx + 5, cy - 5)], fill=_GREEN + (255,), width=3)
    # pedestrians, one colour per status
    positions = np.asarray(positions, dtype=np.float64).reshape(-1, 2)
    statuses = np.asarray(statuses).reshape(-1)
    xs, ys = view.px(positions[:, 0], positions[:, 1])
    for code, color in STATUS_COLORS.items():
        for i in np.nonzero(statuses == code)[0]:
            if np.isfinite(xs[i]) and np.isfinite(ys[i]):
                d.ellipse([xs[i] - 2.5, ys[i] - 2.5, xs[i] + 2.5, ys[i] + 2.5], fill=color + (255,))
    # agent: red +
    gx, gy = (float(v) for v in view.px(ax_, ay_))
    d.line([(gx - 6, gy), (gx + 6, gy)], fill=_RED + (255,), width=2)
    d.line([(gx, gy - 6), (gx, gy + 6)], fill=_RED + (255,), width=2)
    if title:
        for k, line in enumerate(title.split("\n")[:2]):
            d.text((SIZE // 2 - 3 * len(line), 4 + 14 * k), line, fill=(0, 0, 0, 255))
    return img.convert("RGB")


def save_png(path: str, positions, statuses, agent_position, **kw) -> str:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    draw_frame(positions, statuses, agent_position, **kw).save(path, format="PNG")
    return path


def save_gif(path: str, positions: Sequence, statuses: Sequence, agent_positions: Sequence, *, interval_ms: int = 20, **kw) -> str:
    """Animated GIF of a recorded trajectory: frame i shows pedestrians.memory[i] with the agent at agent.memory[i]
    (the reference's `update(i)`, env.py:300-312)."""
    n = min(len(positions), len(statuses), len(agent_positions))
    if n == 0:
        raise RuntimeError("no trajectory recorded: construct the env with draw=True (or set env.unwrapped.draw) before reset()")
    frames = [draw_frame(positions[i], statuses[i], agent_positions[i], **kw) for i in range(n)]
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    frames[0].save(path, format="GIF", save_all=True, append_images=frames[1:], duration=interval_ms, loop=0)
    return path


__all__ = ["draw_frame", "save_png", "save_gif", "STATUS_COLORS", "SIZE"]
