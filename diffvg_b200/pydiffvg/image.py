"""`imwrite` (reference pydiffvg/image.py:6-21) without scikit-image: clip, gamma-encode the colour
channels, quantise to 8 bits and write with Pillow, or -- where Pillow is missing too -- with a
minimal PNG encoder (zlib + CRC from the standard library)."""
import os
import struct
import zlib

import numpy as np

__all__ = ['imwrite']


def _png_bytes(a):
    """uint8 [H, W, C] with C in (1, 2, 3, 4) -> PNG file contents (8-bit, no interlace, filter 0)."""
    h, w, c = a.shape
    color_type = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = b''.join(b'\x00' + a[y].tobytes() for y in range(h))

    def chunk(tag, data):
        body = tag + data
        return struct.pack('>I', len(data)) + body + struct.pack('>I', zlib.crc32(body) & 0xffffffff)

    return (b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', w, h, 8, color_type, 0, 0, 0)) +
            chunk(b'IDAT', zlib.compress(raw, 6)) + chunk(b'IEND', b''))


def imwrite(img, filename, gamma=2.2, normalize=False):
    directory = os.path.dirname(filename)
    if directory != '' and not os.path.exists(directory):
        os.makedirs(directory)
    if not isinstance(img, np.ndarray):
        img = img.data.cpu().numpy()
    img = np.array(img, dtype=np.float32)   # private copy: the reference gamma-encodes the caller's array in place
    if normalize:
        img_rng = np.max(img) - np.min(img)
        if img_rng > 0:
            img = (img - np.min(img)) / img_rng
    img = np.clip(img, 0.0, 1.0)
    if img.ndim == 2:
        img = np.expand_dims(img, 2)
    img[:, :, :3] = np.power(img[:, :, :3], 1.0 / gamma)
    out = (img * 255).astype(np.uint8)
    if filename.lower().endswith('.png'):
        try:
            from PIL import Image
            Image.fromarray(out[:, :, 0] if out.shape[2] == 1 else out).save(filename)
        except ImportError:
            with open(filename, 'wb') as f:
                f.write(_png_bytes(np.ascontiguousarray(out)))
    else:
        from PIL import Image   # other formats need Pillow
        Image.fromarray(out[:, :, 0] if out.shape[2] == 1 else out).save(filename)
