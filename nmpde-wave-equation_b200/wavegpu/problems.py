"""Problem definitions in the reference's parameter-file schema
(src/ParameterReader.cpp:39-126; parameters/*.json of the reference).

The twelve shipped problems are re-stated here as a compact table; `problem()` expands
one into the JSON dictionary the ParameterReader / C-ABI consume, with overrides applied
the way the reference's sweep scripts do it (scripts/scalability_sweep.py:105-121).
`write_json()` emits a parameter file for the main-newmark / main-theta executables."""
import copy
import json

_XYT = "x, y, t"
_XY = "x, y"
_SM = "TT=0.5, XX=0.5, ya=0.333, yb=0.666, k=4.0*pi"
_SM2 = "TT=0.5, XX=0.0, ya=-0.333, yb=0.333, k=4.0*pi"
_SB = "xL=0.15, xR=0.25, yB=0.4, yT=0.6, A=1.0, eps=0.01"

# name: (geometry, nel, T, theta, dt, {block: (constants, expression)})  -- R=1, beta=.25, gamma=.5
_TABLE = {
    "sine-membrane": ("[0.0, 3.0] x [0.0, 1.0]", "180, 60", "60.0", "0.5", "0.05", {
        "G": (_SM, "if(t<=TT && x<XX && y>ya && y<yb, sin(k*t), 0.0)"),
        "DGDT": (_SM, "if(t<=TT && x<XX && y>ya && y<yb, cos(k*t)*k, 0.0)")}),
    "sine-membrane-likedeal2": ("[-1.0, 1.0] x [-1.0, 1.0]", "128", "5.0", "0.5", "0.015625", {
        "G": (_SM2, "if(t<=TT && x<XX && y>ya && y<yb, sin(k*t), 0.0)"),
        "DGDT": (_SM2, "if(t<=TT && x<XX && y>ya && y<yb, cos(k*t)*k, 0.0)")}),
    "standing-mode-wsol": ("[0.0, 1.0] x [0.0, 1.0]", "80", "60.0", "0.0", "0.01", {
        "U0": ("", "sin(pi*x)*sin(pi*y)"),
        "Solution": ("", "cos(sqrt(2)*pi*t)*sin(pi*x)*sin(pi*y)")}),
    "two-modes-wsol": ("[0.0, 1.0] x [0.0, 1.0]", "160", "2.0", "0.5", "0.0035", {
        "U0": ("A1=1.0, A2=0.7", "A1*sin(pi*x)*sin(2*pi*y) + A2*sin(2*pi*x)*sin(pi*y)"),
        "V0": ("A1=1.0, A2=0.7", "0.0"),
        "Solution": ("A1=1.0, A2=0.7",
                     "A1*cos(pi*sqrt(5)*t)*sin(pi*x)*sin(2*pi*y) + A2*cos(pi*sqrt(5)*t)*sin(2*pi*x)*sin(pi*y)")}),
    "five-modes-wsol": ("[0.0, 1.0] x [0.0, 1.0]", "160", "12.0", "0.5", "0.00250", {
        "U0": ("", "0.2*sin(pi*x)*sin(pi*y) + 0.15*sin(2*pi*x)*sin(pi*y) + 0.1*sin(pi*x)*sin(2*pi*y)"
                   " + 0.08*sin(2*pi*x)*sin(2*pi*y) + 0.05*sin(3*pi*x)*sin(pi*y)"),
        "Solution": ("", "0.2*cos(sqrt(2)*pi*t)*sin(pi*x)*sin(pi*y) + 0.15*cos(sqrt(5)*pi*t)*sin(2*pi*x)*sin(pi*y)"
                         " + 0.1*cos(sqrt(5)*pi*t)*sin(pi*x)*sin(2*pi*y)"
                         " + 0.08*cos(2*sqrt(2)*pi*t)*sin(2*pi*x)*sin(2*pi*y)"
                         " + 0.05*cos(sqrt(10)*pi*t)*sin(3*pi*x)*sin(pi*y)")}),
    "dumping-wave": ("[0.0, 1.0] x [0.0, 1.0]", "160", "3.0", "0.5", "0.00350", {
        "F": ("", "exp(-0.1*t)*sin(pi*x)*sin(pi*y)*(0.01*cos(sqrt(2)*pi*t) + 0.8886*sin(sqrt(2)*pi*t))"),
        "U0": ("", "0.2*sin(pi*x)*sin(pi*y)"),
        "Solution": ("", "0.2*exp(-0.1*t)*cos(sqrt(2)*pi*t)*sin(pi*x)*sin(pi*y)")}),
    "gaussian-pulse": ("[0.0, 1.0] x [0.0, 1.0]", "80", "1.2", "0.5", "0.0025", {
        "U0": ("alpha=2000, x0=0.3, y0=0.5", "exp(-alpha*((x-x0)^2 + (y-y0)^2))")}),
    "ricker-wavelet": ("[0.0, 1.0] x [0.0, 1.0]", "100", "2.0", "0.5", "0.0035", {
        "F": ("xs=0.5, ys=0.5, f0=20.0, sigma=0.01",
              "((1 - 2*(pi*f0*(t - 1/f0))^2) * exp(-(pi*f0*(t - 1/f0))^2))"
              " * exp(-((x-xs)^2 + (y-ys)^2) / (2*sigma^2))")}),
    "square-pulsing": ("[0.0, 1.0] x [0.0, 1.0]", "80", "2.0", "0.5", "0.0035", {
        "F": ("xs=0.5, ys=0.5, sigma=0.015, f=5.0",
              "if(sin(2*pi*f*t) > 0, exp(-((x-xs)^2 + (y-ys)^2) / (2*sigma^2)), 0.0)")}),
    "square-bump": ("[0.0, 1.0] x [0.0, 1.0]", "120", "60.0", "0.5", "0.01", {
        "U0": (_SB, "A*0.25*(tanh((x - xL)/eps) - tanh((x - xR)/eps))*(tanh((y - yB)/eps) - tanh((y - yT)/eps))"),
        "V0": (_SB, "-A*0.25*((2/(exp((x - xL)/eps)+exp(-(x - xL)/eps))^2"
                    " - (2/(exp((x - xR)/eps)+exp(-(x - xR)/eps))^2))/eps"
                    " *(tanh((y - yB)/eps) - tanh((y - yT)/eps)))")}),
    "traveling-square-bump": ("[0.0, 3.0] x [0.0, 3.0]", "180, 60", "5.0", "0.5", "0.015625", {
        "U0": ("eps=0.0075, T=0.7, w=0.2, A=1.0", "A*0.5*(tanh(x/eps) - tanh((x - w)/eps))"),
        "V0": ("eps=0.0075, T=0.7, w=0.2, A=1.0, c=1.0",
               "-c*A*0.5*(1/(cosh(x/eps)^2) - 1/(cosh((x - w)/eps)^2))")}),
    "oscillating-boundary": ("[0.0, 1.0] x [0.0, 1.0]", "80", "3.0", "0.5", "0.005", {
        "G": ("", "if(x<0.1 && 0<=y && y<=1, sin(6*pi*t), 0.0)"),
        "DGDT": ("", "if(x<0.1 && 0<=y && y<=1, cos(6*pi*t)*6*pi, 0.0)")}),
}

NAMES = tuple(_TABLE)
_BLOCKS = (("C", _XYT, "1.0"), ("F", _XYT, "0.0"), ("U0", _XY, "0.0"), ("V0", _XY, "0.0"),
           ("G", _XYT, "0.0"), ("DGDT", _XYT, "0.0"))


def problem(name, **overrides):
    """Parameter dictionary for `name`; keyword overrides use the JSON keys with
    spaces replaced by underscores (Nel=..., R=..., Dt=..., Save_Solution=False, C={...})."""
    geom, nel, T, theta, dt, blocks = _TABLE[name]
    p = {"Geometry": geom, "Nel": nel, "R": "1", "T": T, "Theta": theta, "Beta": "0.25",
         "Gamma": "0.5", "Dt": dt}
    for blk, variables, default in _BLOCKS:
        consts, expr = blocks.get(blk, ("", default))
        p[blk] = {"Function constants": consts, "Function expression": expr, "Variable names": variables}
    if "Solution" in blocks:
        consts, expr = blocks["Solution"]
        p["Solution"] = {"Function constants": consts, "Function expression": expr, "Variable names": _XYT}
    for k, v in overrides.items():
        key = k.replace("_", " ")
        p[key] = copy.deepcopy(v) if isinstance(v, dict) else (v if isinstance(v, bool) else str(v))
    return p


def write_json(path, params):
    with open(path, "w") as fh:
        json.dump(params, fh, indent="\t")
