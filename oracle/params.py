"""Parameter-file helpers of the ORACLE (test infrastructure).

Restates src/ParameterReader.cpp:177-230 (get_geometry / get_nel)."""
import re


def parse_geometry(s):
    m = re.fullmatch(
        r"\[\s*([-\d\.]+)\s*,\s*([-\d\.]+)\s*\]\s*x\s*\[\s*([-\d\.]+)\s*,\s*([-\d\.]+)\s*\]", s.strip())
    if not m:
        raise ValueError("Invalid Geometry format in parameters.")
    x0, x1, y0, y1 = (float(g) for g in m.groups())
    return (x0, x1), (y0, y1)


def parse_nel(s):
    toks = [t.strip() for t in str(s).split(",") if t.strip()]
    if len(toks) == 1:
        return int(toks[0]), int(toks[0])
    if len(toks) == 2:
        return int(toks[0]), int(toks[1])
    raise ValueError("Invalid Nel format.")
