"""Minimal NIfTI-1 single-file reader / writer (.nii, .nii.gz).

The reference reads and writes volumes through nibabel (reference mf.py:623-641,
1223-1228), which is an optional dependency there and absent here; this module covers
exactly what the fit path needs: load an n-D array + its affine, save an array with an
affine.  Scaling (scl_slope / scl_inter) is applied on load like nibabel's get_fdata().
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64,
           256: np.int8, 512: np.uint16, 768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).name: k for k, v in _DTYPES.items()}


def _open(path, mode):
    return gzip.open(path, mode) if path.endswith('.gz') else open(path, mode)


def load(path):
    """Returns (data as float64 ndarray, 4x4 affine)."""
    with _open(path, 'rb') as f:
        raw = f.read()
    if len(raw) < 348:
        raise ValueError("%s is not a NIfTI-1 file (too short)" % path)
    endian = '<'
    if struct.unpack('<i', raw[0:4])[0] != 348:
        endian = '>'
        if struct.unpack('>i', raw[0:4])[0] != 348:
            raise ValueError("%s is not a NIfTI-1 file (bad sizeof_hdr)" % path)
    if raw[344:347] not in (b'n+1', b'ni1'):
        raise ValueError("%s is not a NIfTI-1 file (bad magic)" % path)
    dim = struct.unpack(endian + '8h', raw[40:56])
    datatype, = struct.unpack(endian + 'h', raw[70:72])
    pixdim = struct.unpack(endian + '8f', raw[76:108])
    vox_offset, scl_slope, scl_inter = struct.unpack(endian + '3f', raw[108:120])
    qform_code, sform_code = struct.unpack(endian + '2h', raw[252:256])
    if datatype not in _DTYPES:
        raise ValueError("Unsupported NIfTI datatype code %d in %s" % (datatype, path))
    shape = tuple(int(d) for d in dim[1:1 + dim[0]])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(endian)
    n = int(np.prod(shape))
    off = int(vox_offset) if raw[344:347] == b'n+1' else 0
    data = np.frombuffer(raw, dtype=dt, count=n, offset=off).reshape(shape, order='F')
    data = data.astype(np.float64)
    if scl_slope != 0 and not np.isnan(scl_slope) and (scl_slope != 1 or scl_inter != 0):
        data = data * scl_slope + scl_inter
    affine = np.eye(4)
    if sform_code > 0:
        affine[0, :] = struct.unpack(endian + '4f', raw[280:296])
        affine[1, :] = struct.unpack(endian + '4f', raw[296:312])
        affine[2, :] = struct.unpack(endian + '4f', raw[312:328])
    elif qform_code > 0:
        b, c, d, qx, qy, qz = struct.unpack(endian + '6f', raw[256:280])
        a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
        R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                      [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                      [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
        qfac = -1.0 if pixdim[0] < 0 else 1.0
        affine[:3, :3] = R * np.array([pixdim[1], pixdim[2], qfac * pixdim[3]])
        affine[:3, 3] = [qx, qy, qz]
    else:
        affine[:3, :3] = np.diag(pixdim[1:4])
    return np.ascontiguousarray(data), affine


def save(data, affine, path):
    """Writes `data` (any shape up to 7-D) with `affine` as single-file NIfTI-1."""
    data = np.asarray(data)
    if data.dtype.name not in _CODES:
        data = data.astype(np.float64)
    if data.ndim > 7:
        raise ValueError("NIfTI-1 supports at most 7 dimensions")
    affine = np.asarray(affine, dtype=np.float64)
    if affine.shape != (4, 4):
        raise ValueError("affine must have shape (4, 4)")
    hdr = bytearray(348)
    struct.pack_into('<i', hdr, 0, 348)
    dim = [data.ndim] + list(data.shape) + [1] * (7 - data.ndim)
    struct.pack_into('<8h', hdr, 40, *dim)
    struct.pack_into('<h', hdr, 70, _CODES[data.dtype.name])
    struct.pack_into('<h', hdr, 72, data.dtype.itemsize * 8)
    zooms = np.sqrt(np.sum(affine[:3, :3] ** 2, axis=0))
    pixdim = [1.0] + [float(z) for z in zooms] + [1.0] * 4
    struct.pack_into('<8f', hdr, 76, *pixdim)
    struct.pack_into('<f', hdr, 108, 352.0)       # vox_offset
    struct.pack_into('<2f', hdr, 112, 1.0, 0.0)   # scl_slope, scl_inter
    hdr[123] = 10                                  # xyzt_units: mm + s
    struct.pack_into('<2h', hdr, 252, 0, 2)       # qform_code, sform_code (aligned)
    struct.pack_into('<4f', hdr, 280, *affine[0])
    struct.pack_into('<4f', hdr, 296, *affine[1])
    struct.pack_into('<4f', hdr, 312, *affine[2])
    hdr[344:348] = b'n+1\x00'
    with _open(path, 'wb') as f:
        f.write(bytes(hdr))
        f.write(b'\x00\x00\x00\x00')
        f.write(np.asfortranarray(data).astype(data.dtype.newbyteorder('<')).tobytes(order='F'))
    return path
