"""Unity camera conventions needed to drive the path outside Unity.

The reference reads cam.worldToCameraMatrix, cam.projectionMatrix and the engine global
UNITY_MATRIX_P/VP (= GL.GetGPUProjectionMatrix(cam.projectionMatrix, renderIntoTexture: true) on a
D3D-style device); see package/Runtime/GaussianSplatRenderer.cs:586-592,617-620 and
package/Shaders/SplatUtilities.compute:200,236.  All matrices are float32, and are handed to the
C ABI column-major like UnityEngine.Matrix4x4.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def _norm(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v)


def look_rotation(forward, up=(0.0, 1.0, 0.0)) -> np.ndarray:
    """Quaternion.LookRotation as a 3x3 (columns: right, up, forward; left-handed, +z forward)."""
    f = _norm(forward)
    r = _norm(np.cross(np.asarray(up, np.float64), f))
    u = np.cross(f, r)
    return np.stack([r, u, f], axis=1)


def trs(position=(0, 0, 0), rotation=None, scale=(1, 1, 1)) -> np.ndarray:
    """Matrix4x4.TRS with the rotation given as a 3x3."""
    m = np.eye(4, dtype=np.float64)
    rot = np.eye(3) if rotation is None else np.asarray(rotation, np.float64)
    m[:3, :3] = rot * np.asarray(scale, np.float64)[None, :]
    m[:3, 3] = position
    return m


def quat_to_mat(q_xyzw) -> np.ndarray:
    x, y, z, w = [float(v) for v in q_xyzw]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]], np.float64)


def colmajor(m) -> np.ndarray:
    """4x4 [row][col] -> 16 floats, element (r,c) at c*4+r (UnityEngine.Matrix4x4 memory order)."""
    return np.ascontiguousarray(np.asarray(m, np.float32).T).reshape(16)


@dataclass
class Camera:
    position: np.ndarray = field(default_factory=lambda: np.zeros(3))
    rotation: np.ndarray = field(default_factory=lambda: np.eye(3))  # columns right/up/forward
    fieldOfView: float = 60.0      # vertical, degrees
    nearClipPlane: float = 0.3     # GSTestScene.unity:277-279
    farClipPlane: float = 1000.0
    pixelWidth: int = 1200
    pixelHeight: int = 797

    @property
    def aspect(self) -> float:
        return self.pixelWidth / self.pixelHeight

    @property
    def worldToCameraMatrix(self) -> np.ndarray:
        cam_to_world = trs(self.position, self.rotation)
        m = np.linalg.inv(cam_to_world)
        m[2, :] *= -1.0  # camera space looks down -z (OpenGL convention), Unity docs
        return m.astype(np.float32)

    @property
    def projectionMatrix(self) -> np.ndarray:
        f = 1.0 / np.tan(np.radians(self.fieldOfView) * 0.5)
        n, fa = self.nearClipPlane, self.farClipPlane
        return np.array([[f / self.aspect, 0, 0, 0], [0, f, 0, 0], [0, 0, -(fa + n) / (fa - n), -2 * fa * n / (fa - n)],
                         [0, 0, -1, 0]], np.float32)

    def gpuProjectionMatrix(self, renderIntoTexture: bool = True) -> np.ndarray:
        """GL.GetGPUProjectionMatrix on a D3D-style (reversed-Z, y-flipped render texture) device."""
        p = self.projectionMatrix.astype(np.float64).copy()
        if renderIntoTexture:
            p[1, :] *= -1.0
        p[2, :] = p[2, :] * -0.5 + p[3, :] * 0.5
        return p.astype(np.float32)
