"""Volume files of the reference written/read from this engine's buffers (`-F nii|jnii`):

  <session>.nii    NIfTI-1 single file: 348-byte header + 4-byte extender + float64 data, dim = [4, Nx, Ny, Nz, gates]
                   (mcx_savenii, src/mmc_utils.c:515-611; only meaningful for dual-grid output, like the reference)
  <session>.jnii   JNIfTI (JSON + JData): NIFTIHeader + NIFTIData with zlib/base64 payload
                   (mcx_savejnii src/mmc_utils.c:787-905, mcx_jdataencode :995-1076)

Arrays are [gate, z, y, x] in memory (the engine's gate-major volume with x fastest), i.e. NIfTI order x, y, z, t on disk."""
from __future__ import annotations

import base64
import json
import struct
import zlib

import numpy as np

NIFTI_TYPE_FLOAT64 = 64
# nifti_1_header (src/nifti1.h), 348 bytes: sizeof_hdr, data_type, db_name, extents, session_error, regular, dim_info | dim[8] | intent_p1-3 |
# intent_code, datatype, bitpix, slice_start | pixdim[8] | vox_offset, scl_slope, scl_inter | slice_end, slice_code, xyzt_units |
# cal_max, cal_min | slice_duration, toffset | glmax, glmin | descrip, aux_file | qform_code, sform_code | quatern b,c,d, qoffset x,y,z |
# srow_x/y/z | intent_name, magic
_NII = struct.Struct("<i10s18sihBB8h3f4h8f3fhBB2f2f2i80s24s2h6f12f16s4s")
assert _NII.size == 348


def savenii(path, vol, steps=(1.0, 1.0, 1.0), tstep=0.0):
    """vol: float64 [gates, Nz, Ny, Nx] (or [Nz, Ny, Nx]); tstep in seconds (stored in microseconds like the reference)."""
    v = np.ascontiguousarray(vol, dtype=np.float64)
    if v.ndim == 3:
        v = v[None]
    ng, nz, ny, nx = v.shape
    dim = [4, nx, ny, nz, ng, 0, 0, 0]
    pixdim = [0.0, float(steps[0]), float(steps[1]), float(steps[2]), float(tstep) * 1e6, 0.0, 0.0, 0.0]
    hdr = _NII.pack(348, b"", b"", 0, 0, 0, 0, *dim, 0.0, 0.0, 0.0, 0, NIFTI_TYPE_FLOAT64, 64, 0, *pixdim,
                    352.0, 0.0, 0.0, 0, 0, 2 | 24, 0.0, 0.0, 0.0, 0.0, 0, 0, b"", b"", 0, 0, *([0.0] * 6), *([0.0] * 12), b"", b"n+1\0")
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(b"\0\0\0\0")
        f.write(v.tobytes())


def loadnii(path):
    raw = open(path, "rb").read()
    h = _NII.unpack_from(raw, 0)
    if h[0] != 348 or not h[-1].startswith(b"n+1"):
        raise ValueError("not a single-file NIfTI-1 volume")
    dim = h[7:15]
    datatype, bitpix = h[19], h[20]
    pixdim = h[22:30]
    off = int(h[30])
    assert dim[0] == 4
    if datatype != NIFTI_TYPE_FLOAT64 or bitpix != 64:
        raise ValueError("only float64 volumes are written by mmc")
    nx, ny, nz, ng = dim[1:5]
    data = np.frombuffer(raw, dtype=np.float64, count=nx * ny * nz * ng, offset=off).reshape(ng, nz, ny, nx).copy()
    return dict(vol=data, steps=tuple(pixdim[1:4]), tstep=pixdim[4] * 1e-6)


def savejnii(path, vol, steps=(1.0, 1.0, 1.0), tstep=0.0, name="mmc_b200", description="MMC volumetric output", maxgate=None):
    """vol: [gates, Nz, Ny, Nx] float32/float64.  The payload is stored column-major over (x, y, z, t) = the engine's memory order,
    announced through _ArrayOrder_ 'c' like mcx_jdataencode(iscol=1)."""
    v = np.ascontiguousarray(vol)
    if v.ndim == 3:
        v = v[None]
    if v.dtype not in (np.float32, np.float64):
        v = v.astype(np.float64)
    ng, nz, ny, nx = v.shape
    dims = [nx, ny, nz] + ([ng] if ng > 1 else [])
    dtype = "double" if v.dtype == np.float64 else "single"
    hdr = {
        "NIIHeaderSize": 348, "Dim": dims, "Param1": 0, "Param2": 0, "Param3": 0, "Intent": 0, "DataType": dtype,
        "BitDepth": v.dtype.itemsize * 8, "FirstSliceID": 0, "VoxelSize": [float(s) for s in steps] + ([float(tstep)] if ng > 1 else []),
        "Orientation": {"x": "r", "y": "a", "z": "s"}, "ScaleSlope": 1, "ScaleOffset": 0, "LastSliceID": int(maxgate or ng),
        "SliceType": 1, "Unit": {"L": "mm", "T": "s"}, "MaxIntensity": 1, "MinIntensity": 0, "SliceTime": 0, "TimeOffset": 0,
        "Description": description, "AuxFile": "", "QForm": 0, "SForm": 1, "Quatern": {"b": 0, "c": 0, "d": 0},
        "QuaternOffset": {"x": 0, "y": 0, "z": 0}, "Affine": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]], "Name": name, "NIIFormat": "jnifti",
    }
    data = {"_ArrayType_": dtype, "_ArraySize_": dims, "_ArrayOrder_": "c", "_ArrayZipType_": "zlib",
            "_ArrayZipSize_": [1, int(v.size)], "_ArrayZipData_": base64.b64encode(zlib.compress(v.tobytes())).decode()}
    root = {"_DataInfo_": {"JNIFTIVersion": "0.5", "Comment": "Created by mmc_b200", "AnnotationFormat": "https://neurojson.org/jnifti/draft1",
                           "SerialFormat": "https://json.org"}, "NIFTIHeader": hdr, "NIFTIData": data}
    with open(path, "w") as f:
        json.dump(root, f)


def loadjnii(path):
    """Reads the .jnii files written here and by the reference (zlib or uncompressed base64 payload)."""
    root = json.load(open(path))
    hdr, d = root["NIFTIHeader"], root["NIFTIData"]
    dt = {"double": np.float64, "single": np.float32, "uint32": np.uint32}[d["_ArrayType_"]]
    size = list(d["_ArraySize_"])
    if "_ArrayZipData_" in d:
        raw = base64.b64decode(d["_ArrayZipData_"])
        zt = d.get("_ArrayZipType_", "zlib")
        if zt == "zlib":
            raw = zlib.decompress(raw)
        elif zt == "gzip":
            raw = zlib.decompress(raw, 16 + zlib.MAX_WBITS)
        elif zt != "base64":
            raise ValueError("unsupported _ArrayZipType_ " + zt)
        a = np.frombuffer(raw, dtype=dt)
    else:
        a = np.asarray(d["_ArrayData_"], dtype=dt)
    # column-major over (x, y, z[, t]) = reversed C shape
    a = a.reshape(size[::-1]) if d.get("_ArrayOrder_", "c").lower().startswith("c") else a.reshape(size).T
    if a.ndim == 3:
        a = a[None]
    return dict(vol=a.copy(), dim=hdr["Dim"], steps=tuple(hdr["VoxelSize"][:3]), header=hdr)
