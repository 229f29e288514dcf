"""Test helper: write NERSC files from oracle-layout link arrays with numpy only (independent of the product's writer)."""
import numpy as np

from oracle import qcd


def write_nersc(path, U, floating_point="IEEE64BIG", datatype="4D_SU3_GAUGE_3x3", checksum=None, plaquette=None):
    """U: four arrays [T,Z,Y,X,3,3]"""
    dims = U[0].shape[:4][::-1]
    rows = 2 if datatype == "4D_SU3_GAUGE" else 3
    ftype = np.float64 if "64" in floating_point else np.float32
    big = floating_point.endswith("BIG")
    data = np.stack([u.reshape(-1, 3, 3)[:, :rows, :] for u in U], axis=1)  # [site][mu][rows][3]
    native = np.ascontiguousarray(data.astype(np.complex128 if ftype is np.float64 else np.complex64)).view(ftype)
    cs = int(np.frombuffer(native.tobytes(), dtype="<u4").sum(dtype=np.uint64) & 0xFFFFFFFF)
    P = qcd.plaquette(U) if plaquette is None else plaquette
    L = sum(np.trace(u, axis1=-2, axis2=-1).real.mean() for u in U) / 12.0
    header = "\n".join([
        "BEGIN_HEADER", "HDR_VERSION = 1.0", f"DATATYPE = {datatype}", "STORAGE_FORMAT = 1.0",
        f"DIMENSION_1 = {dims[0]}", f"DIMENSION_2 = {dims[1]}", f"DIMENSION_3 = {dims[2]}", f"DIMENSION_4 = {dims[3]}",
        f"LINK_TRACE = {L:.10g}", f"PLAQUETTE = {P:.10g}", f"CHECKSUM = {cs if checksum is None else checksum:x}",
        "ENSEMBLE_ID = test", "ENSEMBLE_LABEL = test", "SEQUENCE_NUMBER = 1", f"FLOATING_POINT = {floating_point}", "END_HEADER", ""])
    with open(path, "wb") as f:
        f.write(header.encode())
        f.write(native.astype(native.dtype.newbyteorder(">" if big else "<")).tobytes())
    return cs
