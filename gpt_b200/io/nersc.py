"""
NERSC gauge configurations: g.load(fn) / g.save(fn, U, g.format.nersc(...)).

Mirror of lib/gpt/core/io/nersc_io.py (header keys, data types 4D_SU3_GAUGE / 4D_SU3_GAUGE_3x3, IEEE32/64 BIG/LITTLE,
checksum, plaquette and link-trace validation with the reference's tolerances :226-247, writer :264-395).  The host side
only parses the header and reads the file; byte order, third-row reconstruction, the [site][mu] -> [mu][site] reordering into
the device layout and the checksum are one CUDA kernel (cgptb_nersc_munge, gpt_b200/csrc/gauge.cu), and the plaquette /
link trace of the loaded field are computed on the device as well.  Single rank only in this version.
"""
import datetime
import getpass
import os
import socket

import numpy as np

import gpt_b200 as g
from gpt_b200 import capi
from gpt_b200.params import params_convention

_FLOAT = {
    "IEEE64BIG": (8, True), "IEEE64LITTLE": (8, False), "IEEE64": (8, False),
    "IEEE32BIG": (4, True), "IEEE32LITTLE": (4, False), "IEEE32": (4, False),
}
_ROWS = {"4D_SU3_GAUGE_3x3": 3, "4D_SU3_GAUGE": 2}


class format:  # noqa: A001  (GPT's name)
    class nersc:
        @params_convention(label="", id="gpt", sequence_number=1)
        def __init__(self, params):
            self.params = params


def read_header(path):
    """(metadata dict, offset of the data) or None if this is not a NERSC file (nersc_io.py:32-63)"""
    if not os.path.isfile(path):
        return None
    md = {}
    with open(path, "rb") as f:
        try:
            line = f.readline().decode("utf-8").strip()
        except UnicodeDecodeError:
            return None
        if line != "BEGIN_HEADER":
            return None
        while True:
            raw = f.readline()
            if not raw:
                return None
            line = raw.decode("utf-8").strip()
            if line == "END_HEADER":
                break
            field, val = line.split("=", 1)
            md[field.strip()] = val.strip()
        return md, f.tell()


def _tolerance(text, eps):
    digits = len(text.split(".")[1].lower().split("e")[0])
    return max(1e2 * eps, 10.0 ** (-digits + 2))


def load(filename, precision=None):
    hdr = read_header(filename)
    if hdr is None:
        raise NotImplementedError(f"{filename} is not a NERSC gauge configuration")
    md, offset = hdr
    dims = [int(md[f"DIMENSION_{i + 1}"]) for i in range(4)]
    if md["FLOATING_POINT"] not in _FLOAT:
        raise NotImplementedError(f"unknown floating point format {md['FLOATING_POINT']}")
    if md["DATATYPE"] not in _ROWS:
        raise NotImplementedError(f"unknown data type {md['DATATYPE']}")
    float_size, big = _FLOAT[md["FLOATING_POINT"]]
    rows = _ROWS[md["DATATYPE"]]
    if precision is None:
        precision = g.double if float_size == 8 else g.single
    gsites = int(np.prod(dims))
    expect = gsites * 4 * rows * 3 * 2 * float_size
    if os.path.getsize(filename) - offset != expect:
        raise RuntimeError(f"{filename}: {os.path.getsize(filename) - offset} bytes of data, header implies {expect}")
    raw = np.fromfile(filename, dtype=np.uint8, offset=offset)
    grid = g.grid(dims, precision)
    U = [g.mcolor(grid) for mu in range(4)]
    cs = capi.nersc_munge(raw, float_size, big, rows, [u.obj for u in U])
    cs_exp = int(md["CHECKSUM"].upper(), 16)
    if cs != cs_exp:
        raise RuntimeError(f"{filename}: checksum {cs:X}, header says {cs_exp:X}")
    # also check plaquette and link trace (nersc_io.py:226-247)
    P, L = capi.gauge_plaquette([u.obj for u in U])
    if abs(P - float(md["PLAQUETTE"])) >= _tolerance(md["PLAQUETTE"], precision.eps):
        raise RuntimeError(f"{filename}: plaquette {P}, header says {md['PLAQUETTE']}")
    if abs(L - float(md["LINK_TRACE"])) >= _tolerance(md["LINK_TRACE"], precision.eps):
        raise RuntimeError(f"{filename}: link trace {L}, header says {md['LINK_TRACE']}")
    for u in U:
        u.metadata = md
    return U


def save(filename, U, fmt=None):
    """4D_SU3_GAUGE_3x3, IEEE64BIG, like the reference's writer (nersc_io.py:264-395); U in double precision"""
    params = (fmt if fmt is not None else format.nersc()).params
    assert len(U) == 4
    grid = U[0].grid
    assert grid.precision is g.double, "single-precision configurations are not written (nersc_io.py:272-274)"
    P, L = capi.gauge_plaquette([u.obj for u in U])
    # [site][mu][3][3] in native order for the checksum, big endian on disk
    data = np.stack([np.asarray(u[:]).reshape(-1, 3, 3) for u in U], axis=1).astype(np.complex128)
    cs = int(np.frombuffer(data.tobytes(), dtype="<u4").sum(dtype=np.uint64) & 0xFFFFFFFF)
    now = datetime.datetime.now(datetime.timezone.utc).strftime("%c %Z")
    header = f"""BEGIN_HEADER
HDR_VERSION = 1.0
DATATYPE = 4D_SU3_GAUGE_3x3
STORAGE_FORMAT =
DIMENSION_1 = {grid.fdimensions[0]}
DIMENSION_2 = {grid.fdimensions[1]}
DIMENSION_3 = {grid.fdimensions[2]}
DIMENSION_4 = {grid.fdimensions[3]}
LINK_TRACE = {L:.15g}
PLAQUETTE  = {P:.15g}
BOUNDARY_1 = PERIODIC
BOUNDARY_2 = PERIODIC
BOUNDARY_3 = PERIODIC
BOUNDARY_4 = PERIODIC
CHECKSUM =   {cs:x}
SCIDAC_CHECKSUMA =          0
SCIDAC_CHECKSUMB =          0
ENSEMBLE_ID = {params['id']}
ENSEMBLE_LABEL = {params['label']}
SEQUENCE_NUMBER = {params['sequence_number']}
CREATOR = {getpass.getuser()}
CREATOR_HARDWARE = {socket.gethostname()}
CREATION_DATE = {now}
ARCHIVE_DATE = {now}
FLOATING_POINT = IEEE64BIG
END_HEADER
"""
    with open(filename, "wb") as f:
        f.write(header.encode("utf-8"))
        f.write(data.view(np.float64).astype(">f8").tobytes())
