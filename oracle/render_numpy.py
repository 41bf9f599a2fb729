"""oracle/render_numpy.py — numpy restatement of /root/reference/src/render.rs.

TEST INFRASTRUCTURE ONLY, parity unpinned (see oracle/lbm_oracle.h): the arithmetic
is ArrayFire's (`mean_all`, `stdev_all`, `sigmoid`, `arg(cplx2)`, `hsv2rgb`, `clamp`),
restated from its documented behaviour: mean/stdev are returned as f64 and cast to
f32 by the caller (src/render.rs:41-46), stdev_all is the population standard
deviation, hsv2rgb is the usual six-sector conversion.
"""
import numpy as np


def _hsv2rgb(h, s, v):
    h6 = h * np.float32(6.0)
    m = h6.astype(np.int32)
    f = h6 - m.astype(np.float32)
    one = np.float32(1.0)
    p = v * (one - s)
    q = v * (one - s * f)
    t = v * (one - s * (one - f))
    r = np.choose(np.clip(m, 0, 6), [v, q, p, p, t, v, v])
    g = np.choose(np.clip(m, 0, 6), [t, v, v, q, p, p, t])
    b = np.choose(np.clip(m, 0, 6), [p, p, t, v, v, q, p])
    return r, g, b


def _to_u8(c):
    """(256.0 * c).round().min(255.0).max(0.0) as u8  (src/render.rs:80-84); Rust rounds half away from zero."""
    x = np.float32(256.0) * c
    return np.clip(np.floor(x + np.float32(0.5)), 0, 255).astype(np.uint8)


def _value(field):
    """src/render.rs:40-47 / :126-133"""
    avg = np.float32(np.mean(field.astype(np.float64)))
    std = np.float32(np.std(field.astype(np.float64)))
    z = (field - avg) * (np.float32(1.0) / std)
    with np.errstate(over="ignore"):
        val = np.float32(1.0) / (np.float32(1.0) + np.exp(-z))
    return np.clip(val.astype(np.float32), 0, 1)


def render_scalar_field(field):
    """src/render.rs:23-89 -> (h, w, 3) uint8"""
    field = field.astype(np.float32)
    val = _value(field)
    hue = np.zeros_like(val)
    sat = np.ones_like(val)
    r, g, b = _hsv2rgb(hue, sat, val)
    return np.stack([_to_u8(r), _to_u8(g), _to_u8(b)], axis=-1)


def render_vector_field(vx, vy):
    """src/render.rs:91-178 -> (h, w, 3) uint8"""
    vx = vx.astype(np.float32)
    vy = vy.astype(np.float32)
    mag = vx * vx + vy * vy
    phase = np.arctan2(vy, vx).astype(np.float32)
    hue = (phase + np.float32(np.pi)) * (np.float32(1.0 / np.pi) * np.float32(0.5))
    hue = np.clip(hue, 0, 1).astype(np.float32)
    sat = np.full_like(mag, np.float32(0.8))
    val = _value(mag)
    r, g, b = _hsv2rgb(hue, sat, val)
    return np.stack([_to_u8(r), _to_u8(g), _to_u8(b)], axis=-1)


def render_geometry(image, solid):
    """src/render.rs:7-21"""
    out = image.copy()
    out[solid.astype(bool)] = (0, 0, 255)
    return out
