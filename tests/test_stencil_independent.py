"""The Shan-Chen stencil table the DEVICE code is compiled with (csrc/ff_stencil.cuh + the isotropy weights of
csrc/lattice.cuh) against an independently written generic rule.

csrc/ff_stencil.cuh is generated from oracle/ff_stencil_tables.h, i.e. device and oracle share one extraction of the
92 / 36 offsets and line-of-sight expressions from lbm_forcing.F90:51-1299, so the GPU-vs-oracle parity tests cannot see
an extraction error there.  tests/textbook_lbm.py::stencil derives the same stencil from first
principles -- every offset whose squared length carries a weight at that isotropy order, and one geometric visibility
rule (a node two or three lattice units away interacts only if a straight or once-bent path to it is fluid) -- and
shares no table with either.  Here the compiled device data (dumped by a host build of the headers,
tests/native/ff_stencil_dump.cpp) must agree with it: same offsets, same weights, and the same ACTIVE / INACTIVE
decision for every entry on random solid patterns.
"""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

import textbook_lbm as tb

HERE = Path(__file__).resolve().parent
SRC = HERE / "native" / "ff_stencil_dump.cpp"
CSRC = HERE.parent / "taxila-lbm_b200" / "csrc"
EXE = HERE / "native" / "_build" / "ff_stencil_dump"


@pytest.fixture(scope="module")
def device_tables():
    EXE.parent.mkdir(exist_ok=True)
    newest = max(p.stat().st_mtime for p in (SRC, CSRC / "ff_stencil.cuh", CSRC / "lattice.cuh"))
    if not EXE.exists() or EXE.stat().st_mtime < newest:
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-o", str(EXE), str(SRC)], check=True)
    return json.loads(subprocess.run([str(EXE)], check=True, capture_output=True, text=True).stdout)


CASES = [("D3Q19", 3, 4), ("D3Q19", 3, 8), ("D2Q9", 2, 4), ("D2Q9", 2, 8), ("D2Q9", 2, 10)]


def device_entries(tab, D, order):
    """{offset: (weight, [alternatives as lists of D-tuples])} of the entries the device uses at this order"""
    out = {}
    for e in tab["entries"]:
        if e["gate"] > order:
            continue
        w = tab["ffw"][str(order)][e["L"]]
        off = tuple(e["off"][:D])
        assert all(v == 0 for v in e["off"][D:])
        assert off not in out, ("duplicate entry", off)
        out[off] = (w, [[tuple(p[:D]) for p in alt] for alt in e["alts"]])
    return out


@pytest.mark.parametrize("name,D,order", CASES)
def test_offsets_and_weights(device_tables, name, D, order):
    dev = device_entries(device_tables[name], D, order)
    gen = {tuple(off): (w, alts) for off, w, alts in tb.stencil(D, order)}
    assert set(dev) == set(gen), (sorted(set(dev) ^ set(gen)))
    for off in gen:
        assert dev[off][0] > 0 and abs(dev[off][0] - gen[off][0]) <= 1e-16 * gen[off][0], (off, dev[off][0], gen[off][0])
        # the entry's shell index is its squared length
    for e in device_tables[name]["entries"]:
        assert e["L"] == sum(v * v for v in e["off"])


def _active(alts, fluid, c, off):
    """the rule both tables encode: target fluid and (no alternative listed, or some alternative's nodes all fluid)"""
    at = lambda o: fluid[tuple(ci + oi for ci, oi in zip(c, o))]  # noqa: E731
    if not at(off):
        return False
    if not alts or any(len(a) == 0 for a in alts):
        return True
    return any(all(at(m) for m in a) for a in alts)


@pytest.mark.parametrize("name,D,order", CASES)
def test_line_of_sight_decisions_on_random_solids(device_tables, name, D, order):
    dev = device_entries(device_tables[name], D, order)
    gen = {tuple(off): alts for off, w, alts in tb.stencil(D, order)}
    rng = np.random.default_rng(100 * D + order)
    c = (3,) * D
    checked = blocked = 0
    for trial in range(1500):
        fluid = rng.random((7,) * D) > rng.choice([0.1, 0.3, 0.5, 0.7])
        fluid[c] = True
        for off in gen:
            a, b = _active(dev[off][1], fluid, c, off), _active(gen[off], fluid, c, off)
            assert a == b, (name, order, off, dev[off][1], gen[off])
            checked += 1
            blocked += fluid[tuple(ci + oi for ci, oi in zip(c, off))] and not a
    assert checked > 1000 and (order == 4 or blocked > 100)  # (order 4 has no line-of-sight rule; the wider ones must bite)


def _oracle_rows(key):
    """[(min_order, L, off, python expression over F(dx,dy,dz))] of oracle/ff_stencil_tables.h (TXO_FF_BLOCKS_D3 / _D2)"""
    import re

    txt = (HERE.parent / "oracle" / "ff_stencil_tables.h").read_text()
    blk = txt[txt.index("#define TXO_FF_BLOCKS_%s(BLOCK)" % key):]
    rows = []
    for line in blk.splitlines()[1:]:
        m = re.match(r"\s*BLOCK\((\d+), (\d+), (-?\d+),(-?\d+),(-?\d+), [^F]*?, (F\(.*\))\)\s*\\?\s*$", line)
        if not m:
            break
        expr = m.group(6).replace("&&", " and ").replace("||", " or ")
        rows.append((int(m.group(1)), int(m.group(2)), tuple(int(m.group(k)) for k in (3, 4, 5)), expr))
    return rows


@pytest.mark.parametrize("name,D,order", CASES)
def test_oracle_table_against_the_generic_rule(name, D, order):
    """the checker's own table (the line-of-sight EXPRESSIONS extracted from lbm_forcing.F90) makes the same decisions
    as the generic rule -- so neither side of the GPU-vs-oracle comparison rests on the extraction alone"""
    rows = [r for r in _oracle_rows("D3" if D == 3 else "D2") if r[0] <= order]
    gen = {tuple(off): alts for off, w, alts in tb.stencil(D, order)}
    assert sorted(r[2][:D] for r in rows) == sorted(gen), (name, order)
    assert all(r[1] == sum(v * v for v in r[2]) for r in rows)
    rng = np.random.default_rng(7 * D + order)
    c = (3,) * D
    for trial in range(600):
        fluid = rng.random((7,) * D) > rng.choice([0.1, 0.3, 0.5, 0.7])
        fluid[c] = True
        F = lambda dx, dy, dz: bool(fluid[tuple(ci + oi for ci, oi in zip(c, (dx, dy, dz)[:D]))])  # noqa: E731
        for mo, L, off, expr in rows:
            assert bool(eval(expr, {"F": F})) == _active(gen[off[:D]], fluid, c, off[:D]), (off, expr, gen[off[:D]])
