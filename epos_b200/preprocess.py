"""Input side of the inference path on the device: resize to max_height_before_crop, crop to the network size and
adjust the intrinsics, as datagen.Dataset._parse_and_preprocess does with TensorFlow ops
(/root/reference/epos_lib/datagen.py:424-476, misc.py:75-91,110-147).  Arithmetic in csrc/cnn_kernels.cu behind
epos_preprocess_u8(); image FILE decoding (tf.image.decode_*) stays on the host."""
import numpy as np
import torch

from . import _lib


def resized_size(in_h, in_w, max_height_before_crop=480):
    new_h = min(max_height_before_crop, in_h)
    return new_h, int(np.float32(in_w) * (np.float32(new_h) / np.float32(in_h)))


def prepare_image(image_u8, K, max_height_before_crop=480, crop_size=(640, 480), offset=None, rng=None, out=None):
    """image_u8: [H,W,3] uint8 CUDA tensor (decoded RGB).  crop_size = (width, height) like the infer_crop_size flag.
    offset (y, x): crop position in the resized image; None draws it uniformly as the reference does (datagen.py:451-455)
    from `rng` (numpy Generator; default: centre... no randomness when the resized image equals the crop).
    Returns (image [crop_h, crop_w, 3] f32 CUDA in [0,255], K' [3,3] f64 numpy)."""
    assert image_u8.is_cuda and image_u8.dtype == torch.uint8 and image_u8.dim() == 3 and image_u8.shape[2] == 3
    image_u8 = image_u8.contiguous()
    H, W = image_u8.shape[:2]
    cw, ch = crop_size
    new_h, new_w = resized_size(H, W, max_height_before_crop)
    if new_h < ch or new_w < cw:
        raise ValueError('resized image %dx%d is smaller than the crop %dx%d' % (new_w, new_h, cw, ch))
    if offset is None:
        rng = rng or np.random.default_rng(0)
        offset = (int(rng.integers(0, new_h - ch + 1)), int(rng.integers(0, new_w - cw + 1)))
    if out is None:
        out = torch.empty((ch, cw, 3), dtype=torch.float32, device=image_u8.device)
    Kin = np.ascontiguousarray(K, np.float64).reshape(9)
    Kout = np.zeros(9, np.float64)
    _lib.check(_lib.lib().epos_preprocess_u8(image_u8.data_ptr(), H, W, image_u8.stride(0), max_height_before_crop, ch, cw,
                                             int(offset[0]), int(offset[1]), Kin.ctypes.data, out.data_ptr(),
                                             Kout.ctypes.data, _lib.stream_ptr()), 'epos_preprocess_u8')
    return out, Kout.reshape(3, 3)
