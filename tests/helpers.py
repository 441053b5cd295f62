"""Shared helpers for the parity tests: record conversion between the oracle's and the engine's patch layouts,
small cached synthetic scenes, and field-by-field comparison."""
from __future__ import annotations

import functools

import numpy as np

import hpmvs_b200 as hp
import oracle


def to_engine(p_or: np.ndarray) -> np.ndarray:
    out = np.zeros(len(p_or), hp.PATCH_DTYPE)
    for f in ("center", "normal", "scale", "nimages"):
        out[f] = p_or[f]
    out["images"] = p_or["images"][:, :hp.MAX_VIEWS]
    return out


def to_oracle(p_en: np.ndarray) -> np.ndarray:
    out = np.zeros(len(p_en), oracle.PATCH_DTYPE)
    for f in ("center", "normal", "scale", "nimages"):
        out[f] = p_en[f]
    out["images"][:, :hp.MAX_VIEWS] = p_en["images"]
    return out


@functools.lru_cache(maxsize=8)
def small_plane(n_views=8, n_seeds=400, seed=2, width=640, height=480, focal=600.0):
    sc = hp.synth.plane_scene(n_views=n_views, width=width, height=height, focal=focal, n_seeds=n_seeds, seed=seed, tex_size=512)
    orc = oracle.OracleScene.from_synth(sc)
    seeds, valid = orc.seed_patches(sc.points, sc.meas_offsets, sc.meas_cam)
    return sc, orc, seeds[valid]


def compare_outputs(ref: np.ndarray, got: np.ndarray):
    """Returns dict of statistics comparing oracle records `ref` with engine records `got`."""
    st = {}
    st["n"] = len(ref)
    st["status_equal"] = int((ref["status"] == got["status"]).sum())
    ok = (ref["status"] == 0) & (got["status"] == 0)
    st["both_ok"] = int(ok.sum())
    same_vis = np.array([r["nimages"] == g["nimages"] and np.array_equal(r["images"][:r["nimages"]], g["images"][:g["nimages"]])
                         for r, g in zip(ref[ok], got[ok])], bool)
    st["vis_equal"] = int(same_vis.sum())
    bit = np.array([np.array_equal(r["center"], g["center"]) and np.array_equal(r["normal"], g["normal"]) and
                    np.array_equal(r["color"], g["color"]) and r["evals"] == g["evals"]
                    for r, g in zip(ref[ok], got[ok])], bool)
    st["bit_exact"] = int(bit.sum())
    if ok.any():
        dc = np.linalg.norm(ref["center"][ok][:, :3] - got["center"][ok][:, :3], axis=1) / ref["scale"][ok]
        dn = np.linalg.norm(ref["normal"][ok][:, :3] - got["normal"][ok][:, :3], axis=1)
        st["max_dcenter_over_scale"] = float(dc.max())
        st["max_dnormal"] = float(dn.max())
        st["max_dscore"] = float(np.abs(ref["last_val"][ok] - got["score"][ok]).max())
        st["dcenter_over_scale"], st["dnormal"], st["dscore"] = dc, dn, np.abs(ref["last_val"][ok] - got["score"][ok])
    return st
