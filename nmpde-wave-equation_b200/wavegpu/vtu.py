"""Reader for the VTU pieces the executables write for `"Save Solution": true`
(host/vtu_writer.cpp: inline base64 arrays, UInt32 length header, no compression)."""
import numpy as np


def read_vtu(path):
    """Minimal reader of the inline-base64 VTU pieces written by host/vtu_writer.cpp:
    returns (points[n,3], connectivity, offsets, types, {name: values})."""
    import base64
    import xml.etree.ElementTree as ET

    dtypes = {"Float32": np.float32, "Float64": np.float64, "Int32": np.int32, "UInt8": np.uint8}
    root = ET.parse(path).getroot()
    assert root.attrib["type"] == "UnstructuredGrid" and root.attrib["header_type"] == "UInt32"
    piece = root.find("UnstructuredGrid/Piece")

    def array(node):
        raw = base64.b64decode(node.text.strip())
        nbytes = int(np.frombuffer(raw[:4], dtype=np.uint32)[0])
        assert len(raw) == 4 + nbytes
        return np.frombuffer(raw[4:], dtype=dtypes[node.attrib["type"]])

    pts = array(piece.find("Points/DataArray")).reshape(-1, 3)
    cells = {a.attrib["Name"]: array(a) for a in piece.findall("Cells/DataArray")}
    data = {a.attrib["Name"]: array(a) for a in piece.findall("PointData/DataArray")}
    assert len(pts) == int(piece.attrib["NumberOfPoints"])
    assert len(cells["offsets"]) == int(piece.attrib["NumberOfCells"])
    return pts, cells["connectivity"], cells["offsets"], cells["types"], data
