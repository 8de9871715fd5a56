"""TEST INFRASTRUCTURE -- CPU restatement of the reference's depth pass (Utils/Render_utils.py:150-178 shaders, :207-266
Renderer, :310-347 render_bust_hair_depth).  NOT part of the product.

**Parity unpinned**: the reference rasterises with OpenGL through moderngl + EGL, neither of which exists in the build
container (nor does a GPU), so no golden could be produced; this file restates what that pass computes from the shader
source: gl_Position = proj * pose * v; colour = -z_cam / depth_range (2.0), interpolated perspective-correctly; depth test
on; clear colour 1; frame flipped on read-back.  Pixel centres at +0.5, inclusive edges (OpenGL's top-left rule only
differs on pixels whose centre lies exactly on an edge)."""
import numpy as np


def render_depth(verts, faces, pose, ndc_prj, H, W):
    """verts [n,3] world, faces [m,3], pose world->camera 4x4, ndc_prj (fx, fy, cx, cy).  -> float64 [H,W]."""
    fx, fy, cx, cy = [float(v) for v in ndc_prj]
    v = np.asarray(verts, dtype=np.float64)
    cam = v @ np.asarray(pose, dtype=np.float64)[:3, :3].T + np.asarray(pose, dtype=np.float64)[:3, 3]
    z = cam[:, 2]
    u = (fx * cam[:, 0] + cx * z) / z
    vv = (fy * cam[:, 1] + cy * z) / z
    px = ((-u) + 1) / 2 * W
    py = (vv + 1) / 2 * H
    out = np.full((H, W), np.inf)
    for f in np.asarray(faces):
        if not np.all(z[f] < -0.1):
            continue
        x, y, iz = px[f], py[f], 1.0 / -z[f]
        area = (x[1] - x[0]) * (y[2] - y[0]) - (x[2] - x[0]) * (y[1] - y[0])
        if area == 0:
            continue
        x0, x1 = max(0, int(np.floor(x.min() - 0.5))), min(W - 1, int(np.ceil(x.max() - 0.5)))
        y0, y1 = max(0, int(np.floor(y.min() - 0.5))), min(H - 1, int(np.ceil(y.max() - 0.5)))
        if x1 < x0 or y1 < y0:
            continue
        qx, qy = np.meshgrid(np.arange(x0, x1 + 1) + 0.5, np.arange(y0, y1 + 1) + 0.5)
        l0 = ((x[1] - qx) * (y[2] - qy) - (x[2] - qx) * (y[1] - qy)) / area
        l1 = ((x[2] - qx) * (y[0] - qy) - (x[0] - qx) * (y[2] - qy)) / area
        l2 = 1 - l0 - l1
        inside = (l0 >= 0) & (l1 >= 0) & (l2 >= 0)
        d = 1.0 / (l0 * iz[0] + l1 * iz[1] + l2 * iz[2])
        sub = out[y0:y1 + 1, x0:x1 + 1]
        sub[inside] = np.minimum(sub[inside], d[inside])
    return np.where(np.isinf(out), 1.0, out / 2.0)
