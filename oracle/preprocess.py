"""ORACLE (test infrastructure, never imported by the product): numpy restatement of the input side of the inference path,
/root/reference/epos_lib/datagen.py:424-476 (_parse_and_preprocess: resize to max_height_before_crop, crop, intrinsics)
with misc.resize_image_tf (misc.py:75-91) = tf.image.resize_area / resize_bilinear with align_corners=True.

Third-party arithmetic not under /root/reference: TensorFlow 1.12 tensorflow/core/kernels/resize_area_op.cc and
resize_bilinear_op.cc, restated from the published kernels (f32 arithmetic).  PARITY UNPINNED for resize_area: the
reference holds no vector for it; the tests check properties (constant images, integer down-scaling = block mean on the
interior, identity at equal size) and the bilinear branch against torch's align_corners=True interpolation."""
import numpy as np


def new_size(in_h, in_w, max_height_before_crop):
    new_h = min(max_height_before_crop, in_h)
    scale = np.float32(new_h) / np.float32(in_h)
    return new_h, int(np.float32(in_w) * scale), scale


def _scale(n_in, n_out):
    return np.float32(n_in - 1) / np.float32(n_out - 1) if n_out > 1 else np.float32(n_in) / np.float32(n_out)


def resize_area(img, new_h, new_w):
    """img [H,W,3] uint8 -> [new_h,new_w,3] f32 (align_corners=True)."""
    H, W = img.shape[:2]
    sy, sx = _scale(H, new_h), _scale(W, new_w)
    src = img.astype(np.float32)

    def weights(n_in, n_out, s):
        out = []
        for y in range(n_out):
            a, b = np.float32(y) * s, np.float32(y + 1) * s
            lo, hi = int(np.floor(a)), int(np.ceil(b))
            ws = []
            for i in range(lo, hi):
                fi = np.float32(i)
                w = (s if fi + 1 > b else fi + 1 - a) if fi < a else (b - fi if fi + 1 > b else np.float32(1.0))
                ws.append((min(max(i, 0), n_in - 1), np.float32(w)))
            out.append(ws)
        return out
    # separable: out = Wy . src . Wx^T / (sy sx), with the weights above as dense [n_out, n_in] matrices (f32)
    def dense(ws, n_in):
        m = np.zeros((len(ws), n_in), np.float32)
        for y, row in enumerate(ws):
            for i, w in row:
                m[y, i] += w
        return m
    Wy, Wx = dense(weights(H, new_h, sy), H), dense(weights(W, new_w, sx), W)
    norm = np.float32(1.0) / (sy * sx)
    out = np.einsum('yi,ijc,xj->yxc', Wy, src, Wx, optimize=True).astype(np.float32) * norm
    return out.astype(np.float32)


def resize_bilinear(img, new_h, new_w):
    H, W = img.shape[:2]
    sy = np.float32(H - 1) / np.float32(new_h - 1) if new_h > 1 else np.float32(0)
    sx = np.float32(W - 1) / np.float32(new_w - 1) if new_w > 1 else np.float32(0)
    src = img.astype(np.float32)
    fy = (np.arange(new_h, dtype=np.float32) * sy)[:, None]
    fx = (np.arange(new_w, dtype=np.float32) * sx)[None, :]
    y0, x0 = np.floor(fy).astype(int), np.floor(fx).astype(int)
    y1, x1 = np.minimum(y0 + 1, H - 1), np.minimum(x0 + 1, W - 1)
    ly, lx = (fy - y0)[..., None], (fx - x0)[..., None]
    top = src[y0, x0] + (src[y0, x1] - src[y0, x0]) * lx
    bot = src[y1, x0] + (src[y1, x1] - src[y1, x0]) * lx
    return (top + (bot - top) * ly).astype(np.float32)


def preprocess(img, K, max_height_before_crop=480, crop=(640, 480), offset=(0, 0)):
    """-> (image [crop_h, crop_w, 3] f32, K' [3,3] f64); crop = (width, height) as infer_crop_size, offset = (y, x)."""
    H, W = img.shape[:2]
    new_h, new_w, s = new_size(H, W, max_height_before_crop)
    r = resize_area(img, new_h, new_w) if H >= new_h else resize_bilinear(img, new_h, new_w)
    oy, ox = offset
    cw, ch = crop
    assert 0 <= oy and oy + ch <= new_h and 0 <= ox and ox + cw <= new_w
    K = np.asarray(K, np.float64)
    fx, fy = np.float32(K[0, 0]) * s, np.float32(K[1, 1]) * s
    cx, cy = np.float32(K[0, 2]) * s - np.float32(ox), np.float32(K[1, 2]) * s - np.float32(oy)
    Ko = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], np.float64)
    return r[oy:oy + ch, ox:ox + cw].copy(), Ko
