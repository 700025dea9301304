"""Read_write (read_write.ml): text format parity with C printf "%g" and the
round-trip expectations of test/read_write_test.ml (tolerance 1e-3 because the
format keeps 6 significant digits).  Host-only: runs without a GPU."""
import os

import numpy as np
import pytest

from mcmc_ocaml_b200 import Failure, InvalidArgument, read_write


def test_text_format_is_printf_g(tmp_path):
    rows = np.array([[0.1, -2.5e-7, 1234567.0, -3.25, 0.0], [1e300, 1e-300, -0.0, np.inf, -np.inf]])
    p = str(tmp_path / "s.dat")
    read_write.write(p, rows)
    lines = open(p).read().splitlines()
    # read_write.ml:19-24: "%g " per coordinate, then "%g %g\n"
    want = [" ".join("%g" % v for v in r) for r in rows]
    assert lines == want
    assert lines[0] == "0.1 -2.5e-07 1.23457e+06 -3.25 0"


def test_read_write_inverses(tmp_path, og):  # read_write_test.ml:36-58
    from mcmc_ocaml_b200 import plugins as P
    out, _, _ = og.mcmc_array(3, 0, 1000, P.gauss_diag([0.0], [1.0]), P.zero(1), P.box_proposal([0.5]), [0.0])
    rows = np.ascontiguousarray(out[:, :, 0])
    p = str(tmp_path / "chain.dat")
    read_write.write(p, rows)
    back = read_write.read(p)
    assert back.shape == rows.shape
    np.testing.assert_allclose(back, rows, rtol=1e-3, atol=1e-3)          # the reference's own tolerance
    read_write.write(p, rows, lossless=True)
    assert np.array_equal(read_write.read(p), rows)                        # "%.17g": exact


def test_nested_read_write(tmp_path):  # read_write_test.ml:60-95
    rng = np.random.default_rng(0)
    rows = np.concatenate([rng.random((50, 3)), np.sort(rng.normal(-5, 2, (50, 1)), axis=0), np.zeros((50, 1))], axis=1)
    lw = rng.normal(-4, 1, 50)
    p = str(tmp_path / "nested.dat")
    read_write.write_nested(p, -1.25, -6.5, rows, lw)
    lev, ldev, r2, lw2 = read_write.read_nested(p)
    assert lev == pytest.approx(-1.25, rel=0.01) and ldev == pytest.approx(-6.5, rel=0.01)
    np.testing.assert_allclose(r2, rows, rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(lw2, lw, rtol=0.01)
    first = open(p).readline()
    assert first == "-1.25 -6.5\n"                                         # read_write.ml:61


def test_errors(tmp_path):
    with pytest.raises(Failure):
        read_write.read(str(tmp_path / "missing.dat"))                     # Sys_error
    p = str(tmp_path / "bad.dat")
    open(p, "w").write("0.1 0.2 0.3\n0.1 oops 0.3\n")
    with pytest.raises(Failure):
        read_write.read(p)                                                 # Scanf failure
    open(p, "w").write("0.1\n")
    with pytest.raises(InvalidArgument):
        read_write.read(p)                                                 # fewer than two fields (Array.sub)
    open(p, "w").write("")
    assert read_write.read(p).shape[0] == 0
