"""Seeded synthetic N-view scenes in the reference's input vocabulary (NVM cameras + points).

The reference consumes a VisualSFM NVM model: cameras ``file f qw qx qy qz cx cy cz r 0`` and points
with per-point measurement lists (/root/reference/src/hpmvs/NVMReader.cpp:31-74).  It ships no sample
data, so the BASELINE.json configs are realised here as procedural scenes: textured planes (and boxes)
rendered by exact inverse ray casting into u8 RGB images, cameras on arcs / loops looking at the scene,
and seed points with measurement lists.  Everything is numpy + a PCG64 seed => reproducible.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------------
# texture
# --------------------------------------------------------------------------------------------
def noise_texture(size: int, seed: int, octaves: int = 6, base: int = 4) -> np.ndarray:
    """Multi-octave value noise, u8 RGB [size, size, 3], mean ~128, sigma ~40."""
    rng = np.random.Generator(np.random.PCG64(seed))
    acc = np.zeros((size, size, 3), np.float64)
    amp_sum = 0.0
    ys = np.arange(size, dtype=np.float64)
    for k in range(octaves):
        g = base << k
        grid = rng.random((g + 1, g + 1, 3))
        t = ys * (g / size)
        i0 = np.minimum(t.astype(np.int64), g - 1)
        fr = t - i0
        # smoothstep interpolation
        fr = fr * fr * (3 - 2 * fr)
        rows = grid[i0] * (1 - fr)[:, None, None] + grid[i0 + 1] * fr[:, None, None]      # [size, g+1, 3]
        tex = rows[:, i0] * (1 - fr)[None, :, None] + rows[:, i0 + 1] * fr[None, :, None]  # [size, size, 3]
        amp = 0.62 ** k
        acc += amp * (tex - 0.5)
        amp_sum += amp
    acc /= np.sqrt((acc ** 2).mean()) + 1e-12
    img = 128.0 + 42.0 * acc
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# cameras
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class NVMCamera:
    """One NVM camera line (NVMReader.h:44-50): focal, rotation quaternion wxyz (world->camera), centre."""
    filename: str
    f: float
    q: np.ndarray   # [4] w x y z
    c: np.ndarray   # [3]
    r: float = 0.0
    width: int = 0
    height: int = 0

    def rotation(self) -> np.ndarray:
        w, x, y, z = self.q
        return np.array([
            [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
            [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
            [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


def quat_from_rotation(R: np.ndarray) -> np.ndarray:
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = math.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = math.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = math.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    q = np.array(q, np.float64)
    return q / np.linalg.norm(q)


def look_at(center: Sequence[float], target: Sequence[float], up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """world->camera rotation with +z forward, +x right, +y down-ish (image y)."""
    c = np.asarray(center, np.float64)
    z = np.asarray(target, np.float64) - c
    z /= np.linalg.norm(z)
    x = np.cross(np.asarray(up, np.float64), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.stack([x, y, z], 0)


# --------------------------------------------------------------------------------------------
# geometry: textured quads (a plane is one big quad; a box is 5 quads)
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class Quad:
    origin: np.ndarray   # [3] corner
    eu: np.ndarray       # [3] edge vector u (full length)
    ev: np.ndarray       # [3] edge vector v (full length)
    tex: np.ndarray      # u8 [T,T,3]

    @property
    def normal(self) -> np.ndarray:
        n = np.cross(self.eu, self.ev)
        return n / np.linalg.norm(n)


def _bilinear(tex: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    T = tex.shape[0]
    x = np.clip(u * (T - 1), 0, T - 1 - 1e-9)
    y = np.clip(v * (T - 1), 0, T - 1 - 1e-9)
    x0 = x.astype(np.int64); y0 = y.astype(np.int64)
    fx = (x - x0)[..., None]; fy = (y - y0)[..., None]
    t = tex.astype(np.float64)
    a = t[y0, x0] * (1 - fx) + t[y0, x0 + 1] * fx
    b = t[y0 + 1, x0] * (1 - fx) + t[y0 + 1, x0 + 1] * fx
    return a * (1 - fy) + b * fy


USE_GPU_RENDERER = False     # bench.py turns this on for the large scenes; tests / goldens always use numpy (bit-reproducible)


def render(cam: NVMCamera, quads: List[Quad], background: int = 30, supersample: int = 1) -> np.ndarray:
    """Exact inverse ray casting of textured quads; nearest hit wins. Returns u8 [H,W,3].
    Uses the GPU (torch, float64) when one is present - the 100-view scenes would take minutes in numpy."""
    try:
        import torch
        if USE_GPU_RENDERER and torch.cuda.is_available() and supersample == 1:
            return _render_torch(cam, quads, background)
    except ImportError:
        pass
    return _render_numpy(cam, quads, background, supersample)


def _render_torch(cam: NVMCamera, quads: List[Quad], background: int) -> np.ndarray:
    import torch
    dev = torch.device("cuda")
    f64 = torch.float64
    W, H = cam.width, cam.height
    R = torch.tensor(cam.rotation(), dtype=f64, device=dev)
    c = torch.tensor(cam.c, dtype=f64, device=dev)
    us = torch.arange(W, dtype=f64, device=dev)
    vs = torch.arange(H, dtype=f64, device=dev)
    vv, uu = torch.meshgrid(vs, us, indexing="ij")
    d_cam = torch.stack([(uu - W / 2.0) / cam.f, (vv - H / 2.0) / cam.f, torch.ones_like(uu)], -1)
    d = d_cam @ R
    out = torch.full((H, W, 3), float(background), dtype=f64, device=dev)
    depth = torch.full((H, W), float("inf"), dtype=f64, device=dev)
    for q in quads:
        eu = torch.tensor(q.eu, dtype=f64, device=dev); ev = torch.tensor(q.ev, dtype=f64, device=dev)
        o = torch.tensor(q.origin, dtype=f64, device=dev)
        n = torch.linalg.cross(eu, ev)
        denom = d @ n
        t = ((o - c) @ n) / denom
        X = c + t[..., None] * d
        rel = X - o
        lu = (rel @ eu) / (eu @ eu)
        lv = (rel @ ev) / (ev @ ev)
        hit = (t > 1e-6) & (lu >= 0) & (lu <= 1) & (lv >= 0) & (lv <= 1) & (t < depth) & torch.isfinite(t)
        if not bool(hit.any()):
            continue
        tex = torch.from_numpy(q.tex).to(dev).to(f64)
        T = tex.shape[0]
        x = torch.clamp(lu[hit] * (T - 1), 0, T - 1 - 1e-9); y = torch.clamp(lv[hit] * (T - 1), 0, T - 1 - 1e-9)
        x0 = x.long(); y0 = y.long()
        fx = (x - x0)[..., None]; fy = (y - y0)[..., None]
        a = tex[y0, x0] * (1 - fx) + tex[y0, x0 + 1] * fx
        b = tex[y0 + 1, x0] * (1 - fx) + tex[y0 + 1, x0 + 1] * fx
        out[hit] = a * (1 - fy) + b * fy
        depth[hit] = t[hit]
    return torch.clamp(torch.round(out), 0, 255).to(torch.uint8).cpu().numpy()


def _render_numpy(cam: NVMCamera, quads: List[Quad], background: int = 30, supersample: int = 1) -> np.ndarray:
    W, H = cam.width, cam.height
    R = cam.rotation()
    ss = supersample
    us = (np.arange(W * ss, dtype=np.float64) + 0.5) / ss - 0.5
    vs = (np.arange(H * ss, dtype=np.float64) + 0.5) / ss - 0.5
    uu, vv = np.meshgrid(us, vs)
    # pixel (u,v) -> camera ray; principal point = image centre (Camera.cpp:40)
    d_cam = np.stack([(uu - W / 2.0) / cam.f, (vv - H / 2.0) / cam.f, np.ones_like(uu)], -1)
    d = d_cam @ R        # R^T applied to each row vector: world direction
    out = np.full((H * ss, W * ss, 3), float(background), np.float64)
    depth = np.full((H * ss, W * ss), np.inf)
    for q in quads:
        n = np.cross(q.eu, q.ev)
        denom = d @ n
        with np.errstate(divide="ignore", invalid="ignore"):
            t = ((q.origin - cam.c) @ n) / denom
        X = cam.c + t[..., None] * d
        rel = X - q.origin
        lu = (rel @ q.eu) / (q.eu @ q.eu)
        lv = (rel @ q.ev) / (q.ev @ q.ev)
        hit = (t > 1e-6) & (lu >= 0) & (lu <= 1) & (lv >= 0) & (lv <= 1) & (t < depth) & np.isfinite(t)
        if not hit.any():
            continue
        col = _bilinear(q.tex, lu[hit], lv[hit])
        out[hit] = col
        depth[hit] = t[hit]
    if ss > 1:
        out = out.reshape(H, ss, W, ss, 3).mean((1, 3))
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# scenes
# --------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SynthScene:
    name: str
    cameras: List[NVMCamera]
    images: List[np.ndarray]                   # u8 [H,W,3] per camera (level 0)
    points: np.ndarray                         # [N,3] float64 seed points (NVM_Point.xyz)
    meas_offsets: np.ndarray                   # [N+1] int32 CSR offsets into meas_cam
    meas_cam: np.ndarray                       # [M] int32 imgIndex per measurement (NVM_Measurement)
    quads: List[Quad] = dataclasses.field(default_factory=list)

    @property
    def n_cameras(self) -> int:
        return len(self.cameras)


def _visible(cam: NVMCamera, X: np.ndarray, margin: float) -> np.ndarray:
    R = cam.rotation()
    pc = (X - cam.c) @ R.T
    z = pc[:, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = cam.f * pc[:, 0] / z + cam.width / 2.0
        v = cam.f * pc[:, 1] / z + cam.height / 2.0
    return (z > 0) & (u >= margin) & (u < cam.width - margin) & (v >= margin) & (v < cam.height - margin)


def plane_scene(n_views: int = 8, width: int = 1280, height: int = 960, focal: float = 1200.0,
                radius: float = 8.0, arc_deg: float = 40.0, n_seeds: int = 10000, extent: float = 2.5,
                seed: int = 2, tex_size: int = 1024, depth_noise: float = 0.5, plane_half: float = 6.0,
                name: Optional[str] = None, elev_deg: float = 6.0, point_seed: Optional[int] = None) -> SynthScene:
    """BASELINE config 2 family: cameras on an arc of `arc_deg` at distance `radius` around a textured
    plane z=0.  Seeds: jittered sqrt(n) x sqrt(n) grid over +-extent, displaced along the plane
    normal by N(0,(depth_noise*scale)^2) with scale = 16*radius/focal (getScale at START_LEVEL 4).
    Every seed is measured in every view that sees it."""
    rng = np.random.Generator(np.random.PCG64(seed if point_seed is None else point_seed))   # seed points only
    tex = noise_texture(tex_size, seed * 7919 + 1)
    quad = Quad(np.array([-plane_half, -plane_half, 0.0]), np.array([2 * plane_half, 0, 0.0]),
                np.array([0, 2 * plane_half, 0.0]), tex)
    cams: List[NVMCamera] = []
    for i in range(n_views):
        a = math.radians(-arc_deg / 2 + arc_deg * (i / max(1, n_views - 1))) if n_views > 1 else 0.0
        e = math.radians(elev_deg * ((i % 3) - 1))
        c = np.array([radius * math.sin(a) * math.cos(e), radius * math.sin(e), -radius * math.cos(a) * math.cos(e)])
        R = look_at(c, (0.0, 0.0, 0.0), up=(0.0, -1.0, 0.0))
        cams.append(NVMCamera(f"view{i:04d}.ppm", focal, quat_from_rotation(R), c, 0.0, width, height))
    images = [render(c, [quad]) for c in cams]
    g = int(math.ceil(math.sqrt(n_seeds)))
    gx, gy = np.meshgrid(np.arange(g), np.arange(g))
    cell = 2 * extent / g
    px = -extent + (gx.ravel() + rng.random(g * g)) * cell
    py = -extent + (gy.ravel() + rng.random(g * g)) * cell
    scale = 16.0 * radius / focal
    pz = rng.normal(0.0, depth_noise * scale, g * g)
    pts = np.stack([px, py, pz], 1)[:n_seeds]
    return _attach_measurements(name or f"plane{n_views}v", cams, images, pts, [quad], margin=40.0)


def _attach_measurements(name, cams, images, pts, quads, margin: float, max_meas: Optional[int] = None,
                         rng: Optional[np.random.Generator] = None) -> SynthScene:
    vis = np.stack([_visible(c, pts, margin) for c in cams], 1)   # [N, C]
    offs = [0]
    mc: List[int] = []
    for i in range(pts.shape[0]):
        ids = np.nonzero(vis[i])[0]
        if max_meas is not None and len(ids) > max_meas:
            start = int(rng.integers(0, len(ids) - max_meas + 1)) if rng is not None else 0
            ids = ids[start:start + max_meas]
        mc.extend(int(v) for v in ids)
        offs.append(len(mc))
    return SynthScene(name, cams, images, pts.astype(np.float64), np.asarray(offs, np.int32),
                      np.asarray(mc, np.int32), quads)


def _measure_visibility(cams, quads, surf, nrm, pick):
    """vis[i,c]: surface point i is inside view c, un-occluded (ray cast against every quad) and within 65 deg of
    its face normal; cosang[i,c] = |cos| of that angle.  Vectorised over (points x quads); torch on the GPU if enabled."""
    n_seeds, n_views = surf.shape[0], len(cams)
    org = np.stack([q.origin for q in quads]); eus = np.stack([q.eu for q in quads]); evs = np.stack([q.ev for q in quads])
    nq = np.cross(eus, evs)
    use_torch = False
    if USE_GPU_RENDERER:
        try:
            import torch
            use_torch = torch.cuda.is_available()
        except ImportError:
            use_torch = False
    vis = np.zeros((n_seeds, n_views), bool)
    cosang = np.zeros((n_seeds, n_views))
    cos_lim = math.cos(math.radians(65.0))
    if use_torch:
        import torch
        dev, f64 = torch.device("cuda"), torch.float64
        T = lambda a: torch.tensor(a, dtype=f64, device=dev)
        surf_t, nrm_t, org_t, eu_t, ev_t, nq_t = T(surf), T(nrm), T(org), T(eus), T(evs), T(nq)
        pick_t = torch.tensor(pick, device=dev)
        qid = torch.arange(len(quads), device=dev)[None, :]
        euu = (eu_t * eu_t).sum(1); evv = (ev_t * ev_t).sum(1)
    for ci, cam in enumerate(cams):
        inside = _visible(cam, surf, 60.0)
        if use_torch:
            c = T(cam.c)
            ray = surf_t - c
            dist = ray.norm(dim=1)
            dirs = ray / dist[:, None]
            cosv = (-(dirs * nrm_t).sum(1)).abs()
            occ = torch.zeros(n_seeds, dtype=torch.bool, device=dev)
            for a in range(0, n_seeds, 20000):
                dd = dirs[a:a + 20000]
                denom = dd @ nq_t.T
                t = (((org_t - c) * nq_t).sum(1))[None, :] / denom
                X = c + t[..., None] * dd[:, None, :]
                rel = X - org_t[None]
                lu = (rel * eu_t[None]).sum(-1) / euu[None]; lv = (rel * ev_t[None]).sum(-1) / evv[None]
                hit = torch.isfinite(t) & (t > 1e-6) & (t < dist[a:a + 20000, None] - 1e-3) & (lu >= 0) & (lu <= 1) & (lv >= 0) & (lv <= 1) & \
                    (pick_t[a:a + 20000, None] != qid)
                occ[a:a + 20000] = hit.any(1)
            occluded = occ.cpu().numpy(); cosv = cosv.cpu().numpy()
        else:
            ray = surf - cam.c
            dist = np.linalg.norm(ray, axis=1)
            dirs = ray / dist[:, None]
            cosv = np.abs(np.einsum("ij,ij->i", -dirs, nrm))
            occluded = np.zeros(n_seeds, bool)
            for a in range(0, n_seeds, 20000):
                dd = dirs[a:a + 20000]
                with np.errstate(divide="ignore", invalid="ignore"):
                    denom = dd @ nq.T
                    t = (((org - cam.c) * nq).sum(1))[None, :] / denom
                    X = cam.c + t[..., None] * dd[:, None, :]
                    rel = X - org[None]
                    lu = (rel * eus[None]).sum(-1) / (eus * eus).sum(1)[None]; lv = (rel * evs[None]).sum(-1) / (evs * evs).sum(1)[None]
                    hit = np.isfinite(t) & (t > 1e-6) & (t < dist[a:a + 20000, None] - 1e-3) & (lu >= 0) & (lu <= 1) & (lv >= 0) & (lv <= 1) & \
                        (pick[a:a + 20000, None] != np.arange(len(quads))[None, :])
                occluded[a:a + 20000] = hit.any(1)
        vis[:, ci] = inside & ~occluded & (cosv > cos_lim)
        cosang[:, ci] = cosv
    return vis, cosang


def box_quads(center, size, seed: int, tex_size: int = 512) -> List[Quad]:
    """5 visible faces (no bottom) of an axis-aligned box, each with its own texture."""
    cx, cy, cz = center
    sx, sy, sz = size
    x0, x1, y0, y1, z0, z1 = cx - sx / 2, cx + sx / 2, cy - sy / 2, cy + sy / 2, cz - sz / 2, cz + sz / 2
    faces = [
        (np.array([x0, y0, z0]), np.array([sx, 0, 0.0]), np.array([0, sy, 0.0])),   # z = z0
        (np.array([x0, y0, z1]), np.array([sx, 0, 0.0]), np.array([0, sy, 0.0])),   # z = z1
        (np.array([x0, y0, z0]), np.array([0, 0, sz]), np.array([0, sy, 0.0])),     # x = x0
        (np.array([x1, y0, z0]), np.array([0, 0, sz]), np.array([0, sy, 0.0])),     # x = x1
        (np.array([x0, y0, z0]), np.array([sx, 0, 0.0]), np.array([0, 0, sz])),     # y = y0 (top, y up is -y)
    ]
    return [Quad(o, eu, ev, noise_texture(tex_size, seed * 101 + k)) for k, (o, eu, ev) in enumerate(faces)]


def city_scene(n_views: int = 100, width: int = 1920, height: int = 1080, focal: float = 1500.0,
               n_boxes: int = 8, n_seeds: int = 100000, seed: int = 4, loop_radius: float = 14.0,
               max_meas: int = 11, name: Optional[str] = None, image_sink=None) -> SynthScene:
    """BASELINE config 4/5 family: cameras on a loop around a block of textured boxes standing on a
    textured ground plane; seeds sampled on the faces (with depth noise) and measured in up to
    `max_meas` consecutive views that see them un-occluded-ish (front-facing test only)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    quads: List[Quad] = []
    ground = Quad(np.array([-20.0, 0.0, -20.0]), np.array([40.0, 0, 0]), np.array([0, 0, 40.0]),
                  noise_texture(1024, seed * 31 + 7))
    quads.append(ground)
    for b in range(n_boxes):
        ang = 2 * math.pi * b / n_boxes + rng.normal(0, 0.1)
        rad = rng.uniform(2.0, 6.0)
        sx, sz = rng.uniform(1.5, 3.0, 2)
        sy = rng.uniform(2.0, 5.0)
        quads += box_quads((rad * math.cos(ang), -sy / 2, rad * math.sin(ang)), (sx, sy, sz), seed * 1000 + b)
    cams: List[NVMCamera] = []
    for i in range(n_views):
        a = 2 * math.pi * i / n_views
        c = np.array([loop_radius * math.cos(a), -2.0 - 1.0 * math.sin(3 * a), loop_radius * math.sin(a)])
        R = look_at(c, (0.0, -1.5, 0.0), up=(0.0, -1.0, 0.0))
        cams.append(NVMCamera(f"view{i:04d}.ppm", focal, quat_from_rotation(R), c, 0.0, width, height))
    if image_sink is None:
        images = [render(c, quads) for c in cams]
    else:
        # streaming form for scenes whose level-0 pixels do not fit host memory (500 views x 4K = 12 GB): every view is rendered,
        # handed to image_sink(index, cameras, image) - which uploads it - and dropped; the returned scene carries no images
        images = []
        for i, c in enumerate(cams):
            image_sink(i, cams, render(c, quads))
    # seeds on box side faces + ground, area-weighted
    areas = np.array([np.linalg.norm(np.cross(q.eu, q.ev)) for q in quads])
    areas[0] *= 0.15
    pick = rng.choice(len(quads), n_seeds, p=areas / areas.sum())
    uv = rng.random((n_seeds, 2)) * 0.9 + 0.05
    org = np.stack([q.origin for q in quads]); eus = np.stack([q.eu for q in quads]); evs = np.stack([q.ev for q in quads])
    surf = org[pick] + uv[:, :1] * eus[pick] + uv[:, 1:] * evs[pick]
    nrm = np.stack([q.normal for q in quads])[pick]
    scale = 16.0 * loop_radius / focal
    pts = surf + nrm * rng.normal(0, 0.3 * scale, n_seeds)[:, None]
    # measurements: un-occluded, front-facing views only (ray cast against every quad), most frontal first
    vis, cosang = _measure_visibility(cams, quads, surf, nrm, pick)
    offs = [0]
    mc: List[int] = []
    for i in range(n_seeds):
        ids = np.nonzero(vis[i])[0]
        if len(ids) > max_meas:
            ids = ids[np.argsort(-cosang[i, ids], kind="stable")[:max_meas]]
        else:
            ids = ids[np.argsort(-cosang[i, ids], kind="stable")]
        mc.extend(int(v) for v in ids)
        offs.append(len(mc))
    return SynthScene(name or f"city{n_views}v", cams, images, pts.astype(np.float64), np.asarray(offs, np.int32),
                      np.asarray(mc, np.int32), quads)


# --------------------------------------------------------------------------------------------
# NVM_V3 text + PPM images (file surface, NVMReader.cpp:115-155)
# --------------------------------------------------------------------------------------------
def write_nvm(scene: SynthScene, path: str, write_images: bool = True) -> None:
    import os
    folder = os.path.dirname(os.path.abspath(path))
    os.makedirs(folder, exist_ok=True)
    with open(path, "w") as fh:
        fh.write("NVM_V3\n\n%d\n" % len(scene.cameras))
        for cam in scene.cameras:
            q, c = cam.q, cam.c
            fh.write("%s %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g 0\n" %
                     (cam.filename, cam.f, q[0], q[1], q[2], q[3], c[0], c[1], c[2], cam.r))
        fh.write("\n%d\n" % scene.points.shape[0])
        for i in range(scene.points.shape[0]):
            a, b = scene.meas_offsets[i], scene.meas_offsets[i + 1]
            p = scene.points[i]
            meas = " ".join("%d %d 0 0" % (scene.meas_cam[k], i) for k in range(a, b))
            fh.write("%.17g %.17g %.17g 128 128 128 %d %s\n" % (p[0], p[1], p[2], b - a, meas))
        fh.write("\n0\n")
    if write_images:
        for cam, img in zip(scene.cameras, scene.images):
            with open(os.path.join(folder, cam.filename), "wb") as fh:
                fh.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
                fh.write(np.ascontiguousarray(img).tobytes())
