"""Stroke scripts of the benchmark configs: lists of dab descriptors that the
CUDA path and the CPU oracle consume identically."""
import numpy as np

from . import capi


def _strength(tool, alpha, pressure=1.0, invert=False):
    return float(capi.host_lib().DUNE_sculpt_brush_strength(int(tool), float(alpha), float(pressure), False, bool(invert), 1.0, 1.0))


def line_points(p0, p1, radius, spacing_pct=10.0, count=None):
    """dab centres along p0 -> p1, spaced spacing_pct % of the brush diameter
    (Brush.spacing, types/types_brush_defaults.h:48) unless `count` fixes their number"""
    p0 = np.asarray(p0, dtype=np.float64)
    p1 = np.asarray(p1, dtype=np.float64)
    if count is None:
        step = 2.0 * radius * spacing_pct / 100.0
        count = max(2, int(np.linalg.norm(p1 - p0) / step) + 1)
    t = np.linspace(0.0, 1.0, count)[:, None]
    return (p0[None, :] * (1.0 - t) + p1[None, :] * t).astype(np.float32)


def c1_draw_stroke(dabs=100, radius=0.15, alpha=0.5):
    """C1: draw, straight line across the +Z face of the cube, r = 0.15, alpha = 0.5, SMOOTH
    falloff, area-normal direction"""
    pts = line_points((-0.8, 0.0, 1.0), (0.8, 0.0, 1.0), radius, count=dabs)
    bs = _strength(capi.TOOL_DRAW, alpha)
    return [capi.make_dab(capi.TOOL_DRAW, p, radius, bstrength=bs, view_normal=(0, 0, 1),
                          flags=capi.DAB_FIRST_STEP if i == 0 else 0) for i, p in enumerate(pts)]


def c2_smooth_stroke(dabs=200, radius=0.2, alpha=0.75, sphere_radius=1.0):
    """C2: smooth, great-circle arc on the icosphere"""
    ang = np.linspace(-0.9, 0.9, dabs)
    pts = np.stack([np.sin(ang), np.zeros_like(ang), np.cos(ang)], axis=-1) * sphere_radius
    bs = _strength(capi.TOOL_SMOOTH, alpha)
    out = []
    for i, p in enumerate(pts.astype(np.float32)):
        n = p / max(np.linalg.norm(p), 1e-20)
        out.append(capi.make_dab(capi.TOOL_SMOOTH, p, radius, bstrength=bs, view_normal=n,
                                 flags=capi.DAB_FIRST_STEP if i == 0 else 0))
    return out


C3_RADII_PCT = (1.0, 2.0, 5.0, 10.0, 20.0, 35.0, 50.0)


def c3_radius_sweep(diag, dabs_per_radius=32, radii_pct=C3_RADII_PCT, alpha=0.5, seed=7, extent=0.9,
                    height=0.05, freq=8.0):
    """C3: draw + normals + bounds per dab on the height-field grid; radius sweep in % of the
    bounding-box diagonal, seeded random centres on the surface"""
    rng = np.random.default_rng(seed)
    bs = _strength(capi.TOOL_DRAW, alpha)
    out = []
    first = True
    for pct in radii_pct:
        r = diag * pct / 100.0
        for _ in range(dabs_per_radius):
            x, y = rng.uniform(-extent, extent, size=2)
            z = height * np.sin(freq * x) * np.cos(freq * y)
            out.append(capi.make_dab(capi.TOOL_DRAW, (x, y, z), r, bstrength=bs, view_normal=(0, 0, 1),
                                     flags=capi.DAB_FIRST_STEP if first else 0))
            first = False
    return out


def c4_tool_stroke(tool, diag, dabs=50, radius_pct=10.0, alpha=0.5, seed=11, height=0.05, freq=8.0):
    """C4: one tool, 50 dabs at r = 10 % of the diagonal along a diagonal line of the grid; grab is a
    single anchored location with growing deltas"""
    r = diag * radius_pct / 100.0
    bs = _strength(tool, alpha)
    out = []
    if tool == capi.TOOL_GRAB:
        loc = (0.1, -0.2, height * np.sin(freq * 0.1) * np.cos(freq * -0.2))
        for i in range(dabs):
            delta = np.array([0.3, 0.1, 0.4]) * r * (i + 1) / dabs
            out.append(capi.make_dab(tool, loc, r, bstrength=bs, view_normal=(0, 0, 1), grab_delta=delta,
                                     flags=capi.DAB_FIRST_STEP if i == 0 else 0))
        return out
    pts = line_points((-0.6, -0.5, 0.0), (0.6, 0.5, 0.0), r, count=dabs)
    prev = None
    for i, p in enumerate(pts):
        p = p.copy()
        p[2] = height * np.sin(freq * p[0]) * np.cos(freq * p[1])
        delta = (0.0, 0.0, 0.0) if prev is None else (p - prev)
        kw = dict(bstrength=bs, view_normal=(0, 0, 1), grab_delta=delta, flags=capi.DAB_FIRST_STEP if i == 0 else 0)
        if tool == capi.TOOL_CLAY_STRIPS:
            kw["flags"] |= capi.DAB_PLANE_TRIM
            kw["tip_roundness"] = 0.18
        out.append(capi.make_dab(tool, p, r, **kw))
        prev = p
    return out


def dab_bytes(tool, U, A, T, M, D, first_A=0, mask=False, automask=False, frontface=False, deg=6.0, iters=1,
              area=False, visited_nodes=0, leaves=0, inner=0):
    """Algorithmic bytes of one dab (BASELINE.md section 4 / SURVEY.md 8d).
    U: unique verts of hit leaves, A: all verts of hit leaves, T: their looptris, M: verts moved,
    D: dirty verts, first_A: verts of leaves first touched (undo snapshot), leaves/inner: boxes written."""
    b = 48 * visited_nodes
    per_u = 12 + (4 if mask else 0) + (4 if automask else 0)
    if tool == capi.TOOL_SMOOTH:
        b += iters * (U * per_u + M * (deg * 16 + 8 + 12))
    else:
        if tool == capi.TOOL_INFLATE or frontface:
            per_u += 12
        b += U * per_u + M * 12
        if area:
            b += U * 12 + M * 12  # the sampling pass: positions of all, normals of the verts inside
    b += first_A * 24
    b += T * 12 + A * 12 + D * 12       # normals (position gather shared with the bounds pass)
    b += 24 * leaves + 72 * inner
    return b
