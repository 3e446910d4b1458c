"""ShapeDNA ``.ev`` text files: drop-ins for ``lapy.io.read_ev`` / ``write_ev`` (reference
lapy/io.py:55-283), the on-disk format of the spectra this package computes.

Layout written by the reference (and by :func:`write_ev`, byte for byte)::

     Refine: 0            header fields, one " Key: value" per line, blank line between groups
     ...
    Eigenvalues:
    { v0 ; v1 ; ... }

    Eigenvectors:
    sizes: n k

    { (column 0, comma separated) ;
    (column 1) ;
    ...
    (column k-1) }

Pure host code (text parsing); nothing here touches the device.
"""

from __future__ import annotations

import re

import numpy as np

_TEXT_FIELDS = ("Creator", "File", "User")
_INT_FIELDS = ("Refine", "Degree", "Dimension", "Elements", "DoF", "NumEW", "EulerChar", "Time(pre)", "Time(calcAB)",
               "Time(calcEW)")  # fmt: skip
_FLOAT_FIELDS = ("Area", "Volume", "BLength")
_RENAMED = {"Time(pre)": "TimePre", "Time(calcAB)": "TimeCalcAB", "Time(calcEW)": "TimeCalcEW"}


def _braced(lines: list[str], start: int) -> tuple[str, int]:
    """Text between the first '{' at or after line ``start`` and its closing '}', and the line after."""
    i = start
    while "{" not in lines[i]:
        i += 1
    parts = []
    while True:
        parts.append(lines[i].strip())
        if "}" in lines[i]:
            break
        i += 1
    return re.sub(r"[{}()]", "", "".join(parts)), i + 1


def read_ev(filename: str) -> dict:
    """Parse an EV file into the dictionary ``compute_shapedna`` returns (plus any header fields).

    Header values are cast like the reference does (io.py:95-123: int, except Area / Volume /
    BLength which are float; Creator / File / User stay text when they are not integers); ``Eigenvectors`` has shape ``EigenvectorsSize`` = (n, k), one column
    per eigenvalue.  Raises ``OSError`` for an unreadable file.
    """
    with open(filename) as f:
        lines = f.read().splitlines()
    d: dict = {}
    i = 0
    while i < len(lines):
        line = lines[i].lstrip()
        key = line.split(":", 1)[0] if ":" in line else None
        if key in _TEXT_FIELDS or key in _INT_FIELDS or key in _FLOAT_FIELDS:
            value = line.split(":", 1)[1].strip()
            if key in _TEXT_FIELDS:
                # the reference casts these to int too (io.py:95-123) and so cannot read back a file
                # whose Creator / File / User is a name; integers stay integers, names stay text
                d[key] = int(value) if re.fullmatch(r"[+-]?\d+", value) else value
            else:
                d[_RENAMED.get(key, key)] = float(value) if key in _FLOAT_FIELDS else int(value)
            i += 1
        elif line.startswith("Eigenvalues"):
            text, i = _braced(lines, i + 1)
            d["Eigenvalues"] = np.array(text.split(";")).astype(float)
        elif line.startswith("Eigenvectors"):
            i += 1
            while not lines[i].strip().startswith("sizes"):
                i += 1
            size = np.array(lines[i].split()[1:]).astype(int)
            d["EigenvectorsSize"] = size
            text, i = _braced(lines, i + 1)
            flat = np.array(text.replace(";", " ").replace(",", " ").split()).astype(float)
            if flat.size == size[0] * size[1]:
                d["Eigenvectors"] = flat.reshape(size[1], size[0]).T  # the file stores columns
            else:
                print(f"[Length of eigenvectors is not {size[0]} times {size[1]}.")
        else:
            i += 1
    return d


def write_ev(filename: str, d: dict) -> None:
    """Write ``d`` (must hold 'Eigenvalues') in the ShapeDNA text format (reference io.py:185-283)."""
    if "Eigenvalues" not in d:
        raise ValueError("ERROR: no Eigenvalues specified")
    out = []

    def group(keys, label=lambda k: k):
        for k in keys:
            if k in d:
                out.append(f" {label(k)}: {d[k]}\n")
        out.append("\n")

    group(("Creator", "File", "User", "Refine", "Degree", "Dimension", "Elements", "DoF", "NumEW"))
    group(("Area", "Volume", "BLength", "EulerChar"))
    times = {"TimePre": "Time(Pre) ", "TimeCalcAB": "Time(calcAB) ", "TimeCalcEW": "Time(calcEW) "}
    for k, lab in times.items():
        if k in d:
            out.append(f" {lab}: {d[k]}\n")
    if all(k in d for k in times):
        out.append(f" Time(total ) : {d['TimePre'] + d['TimeCalcAB'] + d['TimeCalcEW']}\n")
    out.append("\n")
    out.append("Eigenvalues:\n{ " + " ; ".join(map(str, d["Eigenvalues"])) + " }\n\n")
    if "Eigenvectors" in d:
        vec = np.asarray(d["Eigenvectors"])
        out.append("Eigenvectors:\nsizes: " + " ".join(map(str, vec.shape)) + "\n\n{ ")
        cols = ["(" + ",".join(map(str, vec[:, j])) + ")" for j in range(vec.shape[1])]
        out.append(" ;\n".join(cols) + " }\n")
    with open(filename, "w") as f:
        f.write("".join(out))
