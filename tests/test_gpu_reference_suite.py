"""The reference's own test_kmer.py / test_main.py, UNCHANGED, against this package (SURVEY.md section 4 "API-compat gate").

__graft_entry__.build() stages the two files byte for byte under baseline/_ref/ (git-ignored; it travels to the GPU box).
They import `kmer`, `records`, `main` by bare name and spawn `python3 main.py` in the working directory, so pytest runs
them in a subprocess whose cwd is the package directory: every name resolves to the B200 implementation.
Reference tests: /root/reference/src/test_kmer.py:291-633, /root/reference/src/test_main.py:76-328.
"""
import hashlib
import os
import subprocess
import sys

import pytest

from conftest import PKG_DIR, ROOT

REF_TESTS = os.path.join(ROOT, "baseline", "_ref")
pytestmark = pytest.mark.gpu


def _run(name, extra=()):
    path = os.path.join(REF_TESTS, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} is not staged (run __graft_entry__.build() where /root/reference exists)")
    if os.path.isdir("/root/reference/src"):   # in the build container: the staged copy must be the unmodified file
        want = hashlib.sha256(open(os.path.join("/root/reference/src", name), "rb").read()).hexdigest()
        assert hashlib.sha256(open(path, "rb").read()).hexdigest() == want
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", PYTHONPATH=PKG_DIR)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", "--rootdir", REF_TESTS, path, *extra],
                       cwd=PKG_DIR, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    return r.stdout


def test_reference_test_kmer_unchanged():
    out = _run("test_kmer.py")
    assert " passed" in out and "failed" not in out


def test_reference_test_main_unchanged():
    # test_validate_file_writable chmods a directory to 0o400 and expects it to be unwritable: it fails for the
    # reference itself when run as root, which is how the GPU box runs
    deselect = ["-k", "not test_validate_file_writable"] if os.geteuid() == 0 else []
    out = _run("test_main.py", deselect)
    assert " passed" in out and "failed" not in out
