"""Shared scene builders for the tests."""
import numpy as np


def camera(g, width=256, height=256, fov=39.09651, pos=(0.0, 0.5, -6.0), forward=(0.0, 0.0, 1.0)):
    return g.Camera(position=np.array(pos, np.float64), rotation=g.look_rotation(forward), fieldOfView=fov, pixelWidth=width,
                    pixelHeight=height)


def lattice_camera(g, width=256, height=256):
    # cfg1: 1k lattice in [-1,1]^3 seen from -z, axis aligned so the un-jittered z planes tie
    return camera(g, width, height, fov=39.09651, pos=(0.0, 0.0, -4.0))


def one_splat(g, pos=(0, 0, 0), scale=(0.1, 0.1, 0.1), quat=(0, 0, 0, 1), opacity=0.8, dc0=(0.9, 0.5, 0.2), sh=None, n_pad=0,
              quality="VeryHigh"):
    """Asset with one hand-made splat (+ optional far-away padding splats)."""
    import ctypes as C
    from unitygaussiansplatting_b200 import _native as N
    n = 1 + n_pad
    rec = np.zeros((n, 62), np.float32)
    q = np.asarray(quat, np.float32)
    q = q / np.linalg.norm(q)
    packed = np.zeros(4, np.float32)
    N.asset_lib().gsa_pack_smallest3(q.ctypes.data, packed.ctypes.data)
    for i in range(n):
        rec[i, 0:3] = pos if i == 0 else (1000.0 + i, 1000.0, 1000.0)
        rec[i, 6:9] = dc0
        if sh is not None and i == 0:
            rec[i, 9:54] = np.asarray(sh, np.float32).reshape(45)
        rec[i, 54] = opacity
        rec[i, 55:58] = scale
        rec[i, 58:62] = packed
    return g.create_asset(rec, quality)


def view_fields(view):
    """(n,10) uint32 SplatViewData -> dict of float arrays."""
    v = np.ascontiguousarray(view, np.uint32)
    f = v.view(np.float32)
    col = v[:, 8:10]
    def h(x):
        return (x & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
    return {"pos": f[:, 0:4], "axis1": f[:, 4:6], "axis2": f[:, 6:8], "r": h(col[:, 0] >> 16), "g": h(col[:, 0]), "b": h(col[:, 1] >> 16),
            "a": h(col[:, 1])}
