#!/usr/bin/env python3
"""Generate oracle/ff_stencil_tables.h from the reference's fluid-fluid forcing source.

TEST INFRASTRUCTURE ONLY (see the header of taxila_oracle.c and DESIGN.md section 1).

The reference hard-codes the Shan-Chen density-gradient stencil as one
`if (walls(...).eq.0 ...) then ... end if` block per stencil offset
(/root/reference/src/lbm/lbm_forcing.F90:51-958 for D3, :960-1299 for D2).
The ORDER of those blocks fixes the floating-point summation order of the
gradient, and the wall tests in each `if` fix the line-of-sight rule, so the
oracle takes both from the source itself instead of from a re-derivation:
this script parses every block into a row

    BLOCK(min_order, L, dx,dy,dz, cx,cy,cz, wx,wy,wz, <line-of-sight expr>)

min_order : smallest isotropy order that executes the block (4, 8 or 10)
L         : index into ffw(:) (squared length of the offset)
d*        : offset of the neighbour whose density is differenced
c*        : signed integer multiplying ffw(L)*(rho(X+d)-rho(X)) in gradrho(:,dir)
w*        : multiplier of ffw(L) added to weightsum(dir)
expr      : C expression over F(dx,dy,dz) == "walls at that offset .eq. 0"

Only numbers (offsets, integer coefficients, boolean structure) are emitted; no
reference source text is copied.  Run here (the reference tree is not on the GPU
box); the generated header is committed.
"""
import re
import sys
from pathlib import Path

REF = Path("/root/reference/src/lbm/lbm_forcing.F90")
OUT = Path(__file__).resolve().parent / "ff_stencil_tables.h"


def join_continuations(lines):
    out, cur = [], ""
    for ln in lines:
        s = ln.rstrip("\n")
        # strip comments (no strings with '!' in these routines)
        if "!" in s:
            s = s[: s.index("!")]
        s = s.rstrip()
        if not s.strip():
            continue
        if s.endswith("&"):
            cur += s[:-1] + " "
        else:
            cur += s
            out.append(cur.strip())
            cur = ""
    return out


IDX = {"i": 0, "j": 1, "k": 2}


def parse_index(tok):
    tok = tok.replace(" ", "")
    m = re.fullmatch(r"([ijk])([+-]\d+)?", tok)
    assert m, tok
    return IDX[m.group(1)], int(m.group(2) or 0)


def parse_offset(args, ndims):
    parts = args.split(",")
    assert len(parts) == ndims, args
    off = [0, 0, 0]
    for p in parts:
        ax, d = parse_index(p)
        off[ax] = d
    return tuple(off)


def cond_to_c(cond, ndims):
    """Fortran logical expression over walls(..).eq.0 -> C expression over F(dx,dy,dz)."""
    s = cond

    def repl(m):
        off = parse_offset(m.group(1), ndims)
        return "F(%d,%d,%d)" % off

    s = re.sub(r"walls\(([^)]*)\)\s*\.eq\.\s*0(?:\.(?![a-zA-Z]))?", repl, s)
    s = s.replace(".and.", " && ").replace(".or.", " || ")
    s = re.sub(r"\s+", " ", s).strip()
    assert "walls" not in s and ".eq." not in s, s
    return s


DIRS = {"X_DIRECTION": 0, "Y_DIRECTION": 1, "Z_DIRECTION": 2}


def parse_routine(lines, ndims):
    """lines: joined statements of one LBMAddFluidFluidForcesD* routine."""
    rows = []
    order_gate = 4
    depth_gate = []  # stack of ('gate'|'block'|'node')
    cur = None
    for st in lines:
        low = st.lower()
        m = re.match(r"if\s*\(\s*dist%disc%isotropy_order\s*>\s*(\d+)\s*\)\s*then", low)
        if m:
            order_gate = {4: 8, 8: 10}[int(m.group(1))]
            depth_gate.append("gate")
            continue
        m = re.match(r"if\s*\((.*)\)\s*then$", st, flags=re.I)
        if m and "walls(" in m.group(1):
            cond = m.group(1)
            # the outer "this node is fluid" test has offset (0,0,0) only
            c = cond_to_c(cond, ndims)
            if c == "F(0,0,0)":
                depth_gate.append("node")
                continue
            cur = {"gate": order_gate, "cond": c, "g": {}, "w": {}, "L": None, "off": None}
            depth_gate.append("block")
            continue
        if re.match(r"end\s*if", low):
            if not depth_gate:
                continue
            kind = depth_gate.pop()
            if kind == "block":
                rows.append(cur)
                cur = None
            elif kind == "gate":
                order_gate = 4 if order_gate == 8 else 8
                # gates are sequential, never nested: after closing the >4 gate the
                # >8 gate (D2 only) opens explicitly, so the value set here is unused
            continue
        if cur is None:
            continue
        # gradrho statement
        m = re.match(
            r"gradrho\(:,(\w+),[ijk,]+\)\s*=\s*gradrho\(:,(\w+),[ijk,]+\)\s*([+-])\s*(?:(\d+)\.\s*\*\s*)?"
            r"dist%disc%ffw\(\s*(\d+)\s*\)\s*\*\s*\(rho\(:,([^)]*)\)\s*-\s*rho\(:,([^)]*)\)\)",
            st,
        )
        if m:
            d1, d2, sign, mult, L, a, b = m.groups()
            d1, d2 = d1.upper(), d2.upper()  # one block spells x_DIRECTION
            assert d1 == d2
            off = parse_offset(a, ndims)
            assert parse_offset(b, ndims) == (0, 0, 0)
            coef = int(mult or 1) * (1 if sign == "+" else -1)
            L = int(L)
            assert cur["L"] in (None, L)
            cur["L"] = L
            assert cur["off"] in (None, off)
            cur["off"] = off
            assert DIRS[d1] not in cur["g"]
            cur["g"][DIRS[d1]] = coef
            continue
        m = re.match(
            r"weightsum\((\w+)\)\s*=\s*weightsum\((\w+)\)\s*\+\s*dist%disc%ffw\(\s*(\d+)\s*\)\s*(?:\*\s*(\d+)\.)?$",
            st,
        )
        if m:
            d1, d2, L, mult = m.groups()
            d1, d2 = d1.upper(), d2.upper()
            assert d1 == d2 and int(L) == cur["L"]
            assert DIRS[d1] not in cur["w"]
            cur["w"][DIRS[d1]] = int(mult or 1)
            continue
        raise SystemExit("unparsed statement inside block: " + st)
    return rows


def extract(src, name):
    start = next(i for i, l in enumerate(src) if re.match(r"\s*subroutine\s+" + name + r"\b", l))
    end = next(i for i in range(start, len(src)) if re.match(r"\s*end subroutine\s+" + name + r"\b", src[i]))
    return src[start:end]


def main():
    src = REF.read_text().splitlines()
    out = []
    out.append("/* GENERATED by oracle/gen_stencil_tables.py from the reference's")
    out.append(" * src/lbm/lbm_forcing.F90 (D3: lines 51-958, D2: lines 960-1299).")
    out.append(" * TEST INFRASTRUCTURE ONLY.  Rows are in the reference's source order, which is")
    out.append(" * the floating-point summation order of gradrho/weightsum.")
    out.append(" * BLOCK(min_order, L, dx,dy,dz, cx,cy,cz, wx,wy,wz, line_of_sight)          */")
    stats = {}
    for name, ndims, macro in (
        ("LBMAddFluidFluidForcesD3", 3, "TXO_FF_BLOCKS_D3"),
        ("LBMAddFluidFluidForcesD2", 2, "TXO_FF_BLOCKS_D2"),
    ):
        body = join_continuations(extract(src, name))
        rows = parse_routine(body, ndims)
        out.append("#define %s(BLOCK) \\" % macro)
        for r in rows:
            off = r["off"]
            c = [r["g"].get(d, 0) for d in range(3)]
            w = [r["w"].get(d, 0) for d in range(3)]
            # structural sanity (the generic rule): coef = offset component, weight = coef^2
            assert c == list(off), (r, c, off)
            assert w == [x * x for x in off], (r, w)
            assert r["L"] == sum(x * x for x in off), r
            out.append(
                "  BLOCK(%d, %d, %d,%d,%d, %d,%d,%d, %d,%d,%d, %s) \\"
                % (r["gate"], r["L"], *off, *c, *w, r["cond"])
            )
            stats.setdefault((macro, r["gate"]), 0)
            stats[(macro, r["gate"])] += 1
        out.append("  /* end */")
        out.append("")
    OUT.write_text("\n".join(out) + "\n")
    for k, v in sorted(stats.items()):
        print(k, v)


if __name__ == "__main__":
    sys.exit(main())
