"""Small host-side (NumPy) geometry helpers used by the Scene mirror for packing,
unpacking and post-condition bookkeeping.  None of this is on the BA hot path: every
per-detection / per-iteration computation runs in the CUDA library
(``mvus_b200/csrc``).  These helpers only convert the few per-camera quantities that
cross the boundary (rotation matrix <-> Rodrigues vector, K <-> (fx,fy,cx,cy)).

Reference behaviour being mirrored:
  * ``Camera.P2vector`` / ``Camera.vector2P``  (reconstruction/common.py:1113-1144)
    use ``cv2.Rodrigues`` in both directions.
"""
import numpy as np


def rodrigues_to_matrix(rvec):
    """Rotation vector -> 3x3 rotation matrix (cv2.Rodrigues(rvec)[0] semantics,
    reconstruction/common.py:1136,1140)."""
    r = np.asarray(rvec, dtype=np.float64).reshape(3)
    theta = np.sqrt(r @ r)
    if theta < np.finfo(np.float64).eps:
        return np.eye(3)
    k = r / theta
    c, s = np.cos(theta), np.sin(theta)
    kx = np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return c * np.eye(3) + (1.0 - c) * np.outer(k, k) + s * kx


def matrix_to_rodrigues(R):
    """3x3 rotation matrix -> rotation vector (cv2.Rodrigues(R)[0] semantics,
    reconstruction/common.py:1119): the matrix is first projected onto SO(3) by SVD,
    then the axis/angle is read from the skew part (or from the symmetric part near pi)."""
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    U, _, Vt = np.linalg.svd(R)
    R = U @ Vt
    rx = R[2, 1] - R[1, 2]
    ry = R[0, 2] - R[2, 0]
    rz = R[1, 0] - R[0, 1]
    s = np.sqrt((rx * rx + ry * ry + rz * rz) * 0.25)
    c = (R[0, 0] + R[1, 1] + R[2, 2] - 1.0) * 0.5
    c = min(1.0, max(-1.0, c))
    theta = np.arccos(c)
    if s < 1e-5:
        if c > 0:
            return np.zeros(3)
        t = (R[0, 0] + 1.0) * 0.5
        x = np.sqrt(max(t, 0.0))
        t = (R[1, 1] + 1.0) * 0.5
        y = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 1] < 0 else 1.0)
        t = (R[2, 2] + 1.0) * 0.5
        z = np.sqrt(max(t, 0.0)) * (-1.0 if R[0, 2] < 0 else 1.0)
        if abs(x) < abs(y) and abs(x) < abs(z) and (R[1, 2] > 0) != (y * z > 0):
            z = -z
        v = np.array([x, y, z])
        v *= theta / np.sqrt(v @ v)
        return v
    vth = 1.0 / (2.0 * s) * theta
    return np.array([rx, ry, rz]) * vth


def look_at(center, target, up=(0.0, 0.0, 1.0)):
    """World->camera rotation for a camera at ``center`` looking at ``target``
    (camera z forward, x right, y down)."""
    center = np.asarray(center, float)
    z = np.asarray(target, float) - center
    z /= np.linalg.norm(z)
    x = np.cross(z, np.asarray(up, float))
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    return np.vstack((x, y, z))
