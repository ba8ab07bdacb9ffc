"""The .jsonx boundary of the drop-in driver (ldpc-lib_b200/host/settings.cpp) against the reference's own
parser / printer: byte-identical output on the golden cases, and the same refusals."""
import os
import subprocess

import pytest

from conftest import ROOT

TOOL = os.path.join(ROOT, "ldpc-lib_b200", "bin", "jsonx_rt")
CASES = os.path.join(ROOT, "tests", "golden", "jsonx_cases")
EXP = os.path.join(ROOT, "tests", "golden", "jsonx_expected")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(TOOL):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ldpc-lib_b200"), "host"])


def expected_cases():
    out = []
    for f in sorted(os.listdir(EXP)):
        base, _, sel = f[:-len(".jsonx")].partition("__")
        out.append((base + ".jsonx", sel, f))
    return out


SELECT = {"settings": "settings", "settings_more": "settings/more", "top_default": "top_default", "settings_snrs": "settings/snrs",
          "settings_via_inner_defaults": "settings/via_inner_defaults", "results": "results", "": ""}


@pytest.mark.parametrize("case,sel,expected", expected_cases())
def test_roundtrip_is_byte_identical_to_reference(tmp_path, case, sel, expected):
    out = tmp_path / "out.jsonx"
    subprocess.check_call([TOOL, os.path.join(CASES, case), SELECT[sel], str(out)])
    assert out.read_text() == open(os.path.join(EXP, expected)).read()


def test_output_reparses_to_itself(tmp_path):
    a, b = tmp_path / "a.jsonx", tmp_path / "b.jsonx"
    subprocess.check_call([TOOL, os.path.join(CASES, "refs_main.jsonx"), "", str(a)])
    subprocess.check_call([TOOL, str(a), "", str(b)])
    assert a.read_text() == b.read_text()


@pytest.mark.parametrize("case,sel", [("refs_main.jsonx", "settings/fallback_only"), ("refs_main.jsonx", "nothing"),
                                      ("basic.jsonx", "alpha/x"), ("missing_file.jsonx", "")])
def test_refusals(tmp_path, case, sel):
    """die() -> message on stderr, exit status 1 (commons_portable.cpp:181-189)."""
    r = subprocess.run([TOOL, os.path.join(CASES, case), sel, str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.strip()


def test_first_duplicate_key_wins_and_defaults_chain(tmp_path):
    out = tmp_path / "o.jsonx"
    subprocess.check_call([TOOL, os.path.join(CASES, "refs_main.jsonx"), "settings/snrs", str(out)])
    assert out.read_text() == "array { 1.7 }\n"                      # files/input32_16.jsonx:7-8 relies on this
    subprocess.check_call([TOOL, os.path.join(CASES, "refs_main.jsonx"), "settings/shared", str(out)])
    assert out.read_text() == '"from-settings"\n'


def test_sim_inputs_parse(tmp_path):
    """The simulation inputs shipped in configs/ parse, including the `array @"..."` reference to the codes file."""
    out = tmp_path / "o.jsonx"
    for name in ("sim_c1_lms", "sim_c1_tasp"):
        subprocess.check_call([TOOL, os.path.join(ROOT, "configs", name + ".jsonx"), "results", str(out)])
        txt = out.read_text()
        assert "matrix (16 32)" in txt and "_decoder_type" in txt
