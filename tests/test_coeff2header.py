"""Coefficient interchange: header writers byte-identical to the reference's
(fixtures from tests/golden/make_header_golden.py, which also pins them to the reference's own
tests/sig_mean_var.h known-answer file), readers round-tripping into engine-ready arrays."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import sk_dsp_comm_b200.coeff2header as c2h

H = np.load(os.path.join(GOLDEN, "headers.npz"))
F = np.load(os.path.join(GOLDEN, "filters.npz"))


def _text(fn, arg, tmp_path, name="t.h"):
    p = tmp_path / name
    fn(str(p), arg)
    return p.read_text()


def test_fir_header_reference_known_answer(tmp_path):
    """tests/test_coeff2header.py:28-45: 501-sample two-tone sequence."""
    n = np.arange(0, 501)
    x = 3 * np.cos(2 * np.pi * 1000 / 48000 * n) + 2 * np.sin(2 * np.pi * 400 / 48000 * n)
    assert _text(c2h.fir_header, x, tmp_path) == str(H["sig_mean_var"])


@pytest.mark.parametrize("k", ["b101", "b256", "b7", "b1", "b33_remez_bpf"])
def test_fir_headers_byte_identical(k, tmp_path):
    assert _text(c2h.fir_header, F[k], tmp_path) == str(H["fir_" + k])
    assert _text(c2h.fir_fix_header, F[k], tmp_path) == str(H["fix_" + k])


@pytest.mark.parametrize("m", [2, 8, 9, 16, 17])
def test_fir_header_line_wrap_boundaries(m, tmp_path):
    assert _text(c2h.fir_header, F["b101"][:m], tmp_path) == str(H["fir_len%d" % m])
    assert _text(c2h.fir_fix_header, F["b101"][40:40 + m], tmp_path) == str(H["fix_len%d" % m])


@pytest.mark.parametrize("k", ["sos6", "sos_butter5", "sos_sharp_lpf", "sos_tenband"])
def test_sos_headers_byte_identical(k, tmp_path):
    assert _text(c2h.iir_sos_header, F[k], tmp_path) == str(H["sos_" + k])


def test_sos_header_single_stage(tmp_path):
    assert _text(c2h.iir_sos_header, F["sos6"][:1], tmp_path) == str(H["sos_single"])


def test_read_fir_round_trip(tmp_path):
    for k in ("b101", "b256", "b7", "b1"):
        p = tmp_path / (k + ".h")
        c2h.fir_header(str(p), F[k])
        h = c2h.read_fir_header(str(p))
        assert h.shape == F[k].shape and np.abs(h - F[k]).max() <= 0.5e-12      # %15.12f
        c2h.fir_fix_header(str(p), F[k])
        hq = c2h.read_fir_header(str(p))
        assert np.array_equal(hq * 2 ** 15, np.rint(F[k] * 2 ** 15))            # exact Q15
    p = tmp_path / "ref.h"
    p.write_text(str(H["fir_b256"]))                                             # a file the reference wrote
    assert np.abs(c2h.read_fir_header(str(p)) - F["b256"]).max() <= 0.5e-12


def test_read_sos_round_trip(tmp_path):
    for k in ("sos6", "sos_butter5", "sos_tenband"):
        p = tmp_path / (k + ".h")
        p.write_text(str(H["sos_" + k]))
        sos = c2h.read_iir_sos_header(str(p))
        assert sos.shape == F[k].shape and np.all(sos[:, 3] == 1)
        ref = F[k] / F[k][:, 3:4]
        assert np.abs(sos - ref).max() <= 1e-6 * np.abs(ref).max()               # %e keeps 7 digits
        # the parsed cascade is accepted as-is by the drop-in class
        import sk_dsp_comm_b200.multirate_helper as mrh
        assert mrh.multirate_IIR(sos).N_forder == mrh.multirate_IIR(F[k]).N_forder


def test_reader_errors(tmp_path):
    p = tmp_path / "bad.h"
    p.write_text("#define M_FIR 4\nfloat32_t h_FIR[M_FIR] = { 1.0, 2.0, 3.0};\n")
    with pytest.raises(ValueError, match="M_FIR says 4 taps, found 3"):
        c2h.read_fir_header(str(p))
    p.write_text("#define STAGES 1\nfloat32_t ba_coeff[5] = { 1, 2, 3, 4 };\n")
    with pytest.raises(ValueError):
        c2h.read_iir_sos_header(str(p))
    p.write_text("nothing here\n")
    with pytest.raises(ValueError, match="no h_FIR initialiser"):
        c2h.read_fir_header(str(p))
