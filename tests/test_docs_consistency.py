"""Documentation that can rot is checked against the sources (CPU only)."""
import glob
import os
import py_compile
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "soccernet_calibration_sportlight_b200")


def _switches_in_sources():
    names = set()
    for path in glob.glob(os.path.join(PKG, "csrc", "*.cu")) + glob.glob(os.path.join(PKG, "csrc", "*.cuh")):
        names |= set(re.findall(r'getenv\("(CAL_[A-Z0-9_]+)"\)', open(path).read()))
    for path in glob.glob(os.path.join(PKG, "*.py")) + [os.path.join(ROOT, "bench.py")]:
        names |= set(re.findall(r'environ\.get\("(CAL_[A-Z0-9_]+)"', open(path).read()))
    return names


def test_every_environment_switch_is_documented():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = sorted(n for n in _switches_in_sources() if n not in doc)
    assert not missing, f"switches read by the code but absent from INTEGRATION.md: {missing}"


def test_documented_switches_exist():
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    table = doc[doc.index("## Environment switches"):]
    documented = set(re.findall(r"`(CAL_[A-Z0-9_]+)`", table))
    stale = sorted(documented - _switches_in_sources())
    assert not stale, f"switches documented but no longer read anywhere: {stale}"


def test_tools_and_entry_points_compile():
    for path in glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, "bench.py"),
                                                                 os.path.join(ROOT, "__graft_entry__.py")]:
        py_compile.compile(path, doraise=True)


def test_profiles_named_in_readme_exist():
    text = open(os.path.join(ROOT, "profiles", "README.md")).read()
    names = re.findall(r"`((?:r\d[a-z])_[A-Za-z0-9_]+\.(?:json|csv|log|txt))`", text)
    assert names
    missing = [n for n in names if not os.path.exists(os.path.join(ROOT, "profiles", n))]
    assert not missing, missing
