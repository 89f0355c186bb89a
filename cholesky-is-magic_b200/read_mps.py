"""MPS reader (read-mps.lisp) + to-standard-form (standard-form.lisp:18-105) + an MPS writer for the
synthetic generators (BASELINE config 4: "MPS written by generator, read via read-mps").

Behaviour restated from the Lisp, not copied:
  * tokens are split on spaces only; a line starts a section iff its first character is not a space
    (read-mps.lisp:37-41); matching of section names is case-insensitive
  * section order: NAME, [OBJSENSE + one line], ROWS, COLUMNS, RHS (mandatory), [RANGES], [BOUNDS], ENDATA
    (:272-289)
  * ROWS: N rows get indices -1, -2, ...; only the first is the objective, the others are dropped
    (:93-94, 138-141); E/G/L rows are numbered in file order
  * COLUMNS / RHS / RANGES lines carry one or two (row, value) pairs (3 or 5 tokens, :152, :178, :207);
    columns are numbered in order of first appearance
  * BOUNDS: LO UP FX FR MI PL (:229-258); defaults lo = 0, hi = +inf (:321-326)
  * row bounds from type / RHS / RANGES (:295-319)
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .standard_form import StandardForm, Triplets


@dataclass
class RowData:          # row-data (:5-6)
    name: str
    type: str           # '=', '>=', '<='
    rhs: float | None = None
    range: float | None = None
    lb: float | None = None
    ub: float | None = None


@dataclass
class ColData:          # col-data (:8-10)
    name: str
    lb: float | None = None
    ub: float | None = None


@dataclass
class MpsData:          # mps-data (:15-33)
    name: str | None = None
    sense: str | None = None
    rows: dict = field(default_factory=dict)
    row_data: list = field(default_factory=list)
    obj_row: list = field(default_factory=list)       # (column, weight)
    columns: dict = field(default_factory=dict)
    col_data: list = field(default_factory=list)
    triplets: list = field(default_factory=list)      # (col, row, value)


class MpsError(ValueError):
    pass


def tokenize_line(line):
    """tokenize-line (:37-41)."""
    toks = [t for t in line.split(" ") if t]
    return toks, (len(line) > 0 and line[0] != " ")


def mps_float(s):
    """mps-float (:110-116): the Lisp reader with double default format (1e3, 1d3, integers)."""
    try:
        return float(s.replace("d", "e").replace("D", "e"))
    except ValueError as e:
        raise MpsError(f"bad number {s!r}") from e


class _Reader:
    def __init__(self, stream):
        self.lines = iter(stream)

    def get_line(self):
        """get-line (:43-47)."""
        for raw in self.lines:
            line = raw.rstrip("\n").rstrip("\r")
            return tokenize_line(line)
        return None, True

    def get_section_line(self):
        toks, sec = self.get_line()
        if not sec:
            raise MpsError("expected a section header")
        return toks


def _is(header, name):
    return header is not None and len(header) >= 1 and [h.lower() for h in header] == [name]


def read_mps(stream) -> MpsData:
    """read-mps (:272-289)."""
    rd = _Reader(stream)
    mps = MpsData()
    header, sec = rd.get_line()
    if not sec or not header or header[0].lower() != "name" or len(header) < 2:
        raise MpsError("NAME section with a name expected")           # read-name (:57-62)
    mps.name = " ".join(header[1:])
    header = rd.get_section_line()
    if _is(header, "objsense"):                                         # read-sense (:64-77)
        toks, sec = rd.get_line()
        if sec or len(toks) != 1:
            raise MpsError("OBJSENSE needs one data line")
        t = toks[0].lower()
        if t in ("max", "maximize"):
            mps.sense = "max"
        elif t in ("min", "minimize"):
            mps.sense = "min"
        else:
            raise MpsError(f"bad OBJSENSE {toks[0]}")
        header = rd.get_section_line()
    if not _is(header, "rows"):
        raise MpsError("ROWS expected")
    nfree = 0
    while True:                                                         # read-rows (:79-108)
        toks, sec = rd.get_line()
        if sec:
            header = toks
            break
        if len(toks) != 2:
            raise MpsError("ROWS lines have two tokens")
        typ, name = toks[0].upper(), toks[1].lower()
        if name in mps.rows:
            raise MpsError(f"duplicate row {name}")
        if typ == "N":
            nfree += 1
            mps.rows[name] = -nfree
        elif typ in ("E", "G", "L"):
            mps.rows[name] = len(mps.row_data)
            mps.row_data.append(RowData(toks[1], {"E": "=", "G": ">=", "L": "<="}[typ]))
        else:
            raise MpsError(f"bad row type {toks[0]}")
    if not _is(header, "columns"):
        raise MpsError("COLUMNS expected")

    def insert_triplet(col, row, val):                                  # :135-150
        key = col.lower()
        if key not in mps.columns:
            mps.columns[key] = len(mps.col_data)
            mps.col_data.append(ColData(col))
        c = mps.columns[key]
        r = mps.rows.get(row.lower())
        if r is None:
            raise MpsError(f"Unknown row {row}")
        v = mps_float(val)
        if r < -1:
            return
        if r == -1:
            mps.obj_row.append((c, v))
        else:
            mps.triplets.append((c, r, v))

    while True:                                                         # read-columns (:118-157)
        toks, sec = rd.get_line()
        if sec:
            header = toks
            break
        if len(toks) not in (3, 5):
            raise MpsError("COLUMNS lines have 3 or 5 tokens")
        insert_triplet(toks[0], toks[1], toks[2])
        if len(toks) == 5:
            insert_triplet(toks[0], toks[3], toks[4])
    if not _is(header, "rhs"):
        raise MpsError("RHS expected (mandatory in the reference's reader)")

    def pairs_section(setter):
        set_name = None
        while True:
            toks, sec = rd.get_line()
            if sec:
                return toks
            if len(toks) not in (3, 5):
                raise MpsError("RHS/RANGES lines have 3 or 5 tokens")
            if set_name is None:
                set_name = toks[0].lower()
            elif set_name != toks[0].lower():
                raise MpsError("more than one RHS/RANGES set")
            setter(toks[1], toks[2])
            if len(toks) == 5:
                setter(toks[3], toks[4])

    def add_rhs(row, val):                                              # read-rhs (:159-186)
        r = mps.rows.get(row.lower())
        if r is None:
            raise MpsError(f"Unknown row {row}")
        v = mps_float(val)
        if r < 0:
            return                       # objective-row RHS is printed and skipped (:167-168)
        if mps.row_data[r].rhs is not None:
            raise MpsError(f"duplicate RHS for {row}")
        mps.row_data[r].rhs = v

    def add_range(row, val):                                            # read-ranges (:188-215)
        r = mps.rows.get(row.lower())
        if r is None:
            raise MpsError(f"Unknown row {row}")
        v = mps_float(val)
        if r < 0:
            return
        if mps.row_data[r].range is not None:
            raise MpsError(f"duplicate RANGE for {row}")
        mps.row_data[r].range = v

    header = pairs_section(add_rhs)
    if _is(header, "ranges"):
        header = pairs_section(add_range)
    if _is(header, "bounds"):                                           # read-bounds (:217-270)
        bound_name = None
        while True:
            toks, sec = rd.get_line()
            if sec:
                header = toks
                break
            if len(toks) not in (3, 4):
                raise MpsError("BOUNDS lines have 3 or 4 tokens")
            typ, bname, col = toks[0].upper(), toks[1].lower(), toks[2].lower()
            val = mps_float(toks[3]) if len(toks) == 4 else None
            if bound_name is None:
                bound_name = bname
            elif bound_name != bname:
                raise MpsError("more than one BOUNDS set")
            if col not in mps.columns:
                raise MpsError(f"Unknown column {toks[2]}")
            d = mps.col_data[mps.columns[col]]
            if typ in ("LO", "UP", "FX") and val is None:
                raise MpsError(f"{typ} bound needs a value")
            if typ == "LO":
                d.lb = val
            elif typ == "UP":
                d.ub = val
            elif typ == "FX":
                d.lb = d.ub = val
            elif typ == "FR":
                d.lb, d.ub = -math.inf, math.inf
            elif typ == "MI":
                d.lb, d.ub = -math.inf, 0.0
            elif typ == "PL":
                d.lb, d.ub = 0.0, math.inf
            else:
                raise MpsError(f"unsupported bound type {toks[0]}")
    if not _is(header, "endata"):
        raise MpsError("ENDATA expected")
    return mps


def read_mps_file(path) -> MpsData:
    """read-mps-file (:291-293)."""
    with open(path, encoding="utf-8") as f:
        return read_mps(f)


def post_process_mps(mps: MpsData) -> MpsData:
    """post-process-mps (:295-326): default sense, row bounds from type/RHS/RANGES, default column bounds."""
    if mps.sense is None:
        mps.sense = "min"
    for row in mps.row_data:
        rhs = 0.0 if row.rhs is None else row.rhs
        rng = row.range
        if rng is not None:
            a = abs(rng)
            if row.type == "<=":
                row.lb, row.ub = rhs - a, rhs
            elif row.type == ">=":
                row.lb, row.ub = rhs, rhs + a
            else:
                row.lb, row.ub = (rhs + rng, rhs) if rng < 0 else (rhs, rhs + rng)
        else:
            if row.type == "<=":
                row.lb, row.ub = -math.inf, rhs
            elif row.type == ">=":
                row.lb, row.ub = rhs, math.inf
            else:
                row.lb, row.ub = rhs, rhs
        assert row.lb <= row.ub
    for col in mps.col_data:
        if col.lb is None:
            col.lb = 0.0
        if col.ub is None:
            col.ub = math.inf
    return mps


def to_standard_form(mps: MpsData) -> StandardForm:
    """to-standard-form (standard-form.lisp:18-105): equality rows keep b = rhs; >= rows get a slack
    column with coefficient -1, <= rows +1, ranged rows +1 with the slack in [0, ub - lb] and b = ub;
    the objective is negated for max problems."""
    post_process_mps(mps)
    rows = [t[1] for t in mps.triplets]
    cols = [t[0] for t in mps.triplets]
    vals = [t[2] for t in mps.triplets]
    nvars0 = len(mps.col_data)
    l = [c.lb for c in mps.col_data]
    u = [c.ub for c in mps.col_data]
    b, types = [], []

    def artificial_var(row, coef, lb=0.0, ub=math.inf):
        assert lb <= ub
        n = len(l)
        l.append(float(lb))
        u.append(float(ub))
        rows.append(row)
        cols.append(n)
        vals.append(float(coef))

    for i, row in enumerate(mps.row_data):
        if row.lb == row.ub:
            types.append(None)
            b.append(row.lb)
        elif row.ub == math.inf:
            b.append(row.lb)
            types.append(">")
            artificial_var(i, -1)
        elif row.lb == -math.inf:
            b.append(row.ub)
            types.append("<")
            artificial_var(i, 1)
        else:
            b.append(row.ub)
            types.append(None)
            artificial_var(i, 1, 0, row.ub - row.lb)
    c = sorted(mps.obj_row, key=lambda p: p[0])
    if mps.sense != "min":
        c = [(i, -v) for i, v in c]
    return StandardForm(
        nvars=len(l), ncons=len(b), c=c,
        A=Triplets(np.asarray(rows, dtype=np.int32), np.asarray(cols, dtype=np.int32),
                   np.asarray(vals, dtype=np.float64)),
        b=np.asarray(b, dtype=np.float64), l=np.asarray(l, dtype=np.float64), u=np.asarray(u, dtype=np.float64),
        type=types, initial_vars=nvars0)


def write_mps(path, sf: StandardForm, name="SYNTH"):
    """Writes a standard-form LP (equality rows, column bounds) in the dialect read_mps accepts:
    space-separated tokens, data lines starting with a space, one (row, value) pair per line."""
    cvec = dict(sf.c)
    if sf.A is None:
        r, cidx = np.nonzero(sf.A_dense)
        trip = Triplets(r.astype(np.int32), cidx.astype(np.int32), sf.A_dense[r, cidx])
    else:
        trip = sf.A
    order = np.lexsort((trip.row, trip.col))
    with open(path, "w", encoding="utf-8") as f:
        f.write(f"NAME {name}\n")
        f.write("ROWS\n N COST\n")
        for i in range(sf.ncons):
            f.write(f" E R{i}\n")
        f.write("COLUMNS\n")
        k = 0
        tr, tc, tv = trip.row[order], trip.col[order], trip.value[order]
        for j in range(sf.nvars):
            wrote = False
            if j in cvec and cvec[j] != 0.0:
                f.write(f" C{j} COST {cvec[j]!r}\n")
                wrote = True
            while k < len(tv) and tc[k] == j:
                f.write(f" C{j} R{tr[k]} {float(tv[k])!r}\n")
                wrote = True
                k += 1
            if not wrote:
                f.write(f" C{j} COST 0.0\n")     # keep first-appearance column numbering intact
        f.write("RHS\n")
        for i in range(sf.ncons):
            f.write(f" RHS R{i} {float(sf.b[i])!r}\n")
        bounds = []
        for j in range(sf.nvars):
            lo, hi = float(sf.l[j]), float(sf.u[j])
            if lo == -math.inf and hi == math.inf:
                bounds.append(f" FR BND C{j}\n")
            elif lo == -math.inf:
                bounds.append(f" MI BND C{j}\n")
                if hi != 0.0:
                    bounds.append(f" UP BND C{j} {hi!r}\n")
            else:
                if lo != 0.0:
                    bounds.append(f" LO BND C{j} {lo!r}\n")
                if hi != math.inf:
                    bounds.append(f" UP BND C{j} {hi!r}\n")
        if bounds:
            f.write("BOUNDS\n")
            f.writelines(bounds)
        f.write("ENDATA\n")
