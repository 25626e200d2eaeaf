"""CPU: the product's Atmosphere.dat loader (csrc/atmosphere.cc, through the host build) against the oracle's literal
emulation of the reference's stream parsing, on the shipped file and on edited copies of it."""
import os

import numpy as np
import pytest

from conftest import ATMOSPHERE


def _load_both(hostsim, path):
    from oracle.ref import Oracle
    rc = hostsim.lib.sim_load(path.encode(), 0)
    if rc != 0:
        return rc, None, None
    m = hostsim.medium()
    o = Oracle(path).constants()
    return 0, m, o


def _restore(hostsim):
    assert hostsim.lib.sim_load(ATMOSPHERE.encode(), 0) == 0


def _assert_same(m, o):
    assert m["max_layers"] == o["max_layers"]
    assert m["B_air"] == o["B_air"] and m["C_air"] == o["C_air"]
    assert m["atmlay_cm"][:4] == o["atmlay_cm"][:4]


def test_shipped_file(hostsim, oracle_built):
    rc, m, o = _load_both(hostsim, ATMOSPHERE)
    assert rc == 0
    _assert_same(m, o)
    assert m["max_layers"] == 4


def test_without_trailing_newline_the_last_row_is_lost_like_in_the_reference(hostsim, oracle_built, tmp_path):
    text = open(ATMOSPHERE).read().rstrip("\n")
    p = tmp_path / "Atmosphere.dat"
    p.write_text(text)
    from oracle.ref import Oracle
    try:
        rc, m, o = _load_both(hostsim, str(p))
        assert rc == 0
        _assert_same(m, o)
        assert Oracle(str(p)).constants()["npoints"] == Oracle(ATMOSPHERE).constants()["npoints"] - 1
    finally:
        _restore(hostsim)


def test_truncated_table_and_shifted_layers(hostsim, oracle_built, tmp_path):
    lines = open(ATMOSPHERE).read().split("\n")
    try:
        # (a) tabulated profile cut at ~6 km: only two per-layer vectors -> MaxLayers 3
        p = tmp_path / "a" / "Atmosphere.dat"
        p.parent.mkdir()
        p.write_text("\n".join(lines[:6] + lines[6:6006]) + "\n")
        rc, m, o = _load_both(hostsim, str(p))
        assert rc == 0
        _assert_same(m, o)
        assert m["max_layers"] == 3
        # (b) layer edges moved (ATMLAY line rewritten): 2 km / 5 km / 12 km
        hdr = lines[:6]
        hdr[1] = " 0.00000000E+00  2.00000000E+05  5.00000000E+05  1.20000000E+06  1.00000000E+07"
        p = tmp_path / "b" / "Atmosphere.dat"
        p.parent.mkdir()
        p.write_text("\n".join(hdr + lines[6:]))
        rc, m, o = _load_both(hostsim, str(p))
        assert rc == 0
        _assert_same(m, o)
        assert m["max_layers"] == 5   # data now reaches above the fourth edge
        # the whole path on this differently layered atmosphere: forward cells and solves against the oracle
        from conftest import assert_forward_close, assert_solve_close
        from oracle.ref import Oracle
        ob = Oracle(str(p))
        rng = np.random.default_rng(17)
        th, h = rng.uniform(90.1, 180, 4000), rng.uniform(1001, 100000, 4000)
        assert_forward_close(hostsim.forward(th, h, 1000.0, -100.0), ob.forward_batch(th, h, 1000.0, -100.0), "shifted layers")
        ang = rng.uniform(91, 179.8, 4000)
        d = (h - 1000 + 100) * np.tan((180 - ang) * 3.1415927 / 180)
        ok_r, ref = ob.solve_cm_batch(h * 100, d * 100, -10000.0, 100000.0)
        ok, out, _ = hostsim.solve_cm(h * 100, d * 100, -10000.0, 100000.0)
        assert_solve_close(ok, out, ok_r, ref, 3.1415927, "shifted layers")
    finally:
        _restore(hostsim)


def test_malformed_files_are_rejected(hostsim, tmp_path):
    try:
        p = tmp_path / "Atmosphere.dat"
        p.write_text("# header only\n 1 2 3\n")
        assert hostsim.lib.sim_load(str(p).encode(), 0) != 0
        assert hostsim.lib.sim_load(str(tmp_path / "missing.dat").encode(), 0) != 0
        lines = open(ATMOSPHERE).read().split("\n")
        lines[2] = " not numbers at all"
        p.write_text("\n".join(lines))
        assert hostsim.lib.sim_load(str(p).encode(), 0) != 0
    finally:
        _restore(hostsim)
