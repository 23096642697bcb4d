#!/usr/bin/env python
"""
ORACLE TEST INFRASTRUCTURE -- not product code.

Runs the UNMODIFIED reference test-suite (``/root/reference/tests``) on the
dependency shims in ``oracle/shims`` (simpy 3.0.11 engine, gym 0.12.5 surface,
pygame/ode import stubs, pytest_mock).  Passing it is what qualifies the shim
engine as the oracle's event-order authority.

The reference tree is read-only, so the tests are executed from a scratch copy
under /tmp (pytest wants to write ``pytest-logs.txt`` and a cache next to them);
the ``gymwipe`` package itself is imported straight from ``/root/reference``.

Only meaningful in the build container: ``/root/reference`` does not exist on
the GPU box.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GYMWIPE_REFERENCE", "/root/reference")


def main() -> int:
    if not os.path.isdir(os.path.join(REF, "gymwipe")):
        print("reference not available at", REF)
        return 2
    scratch = tempfile.mkdtemp(prefix="gymwipe_ref_tests_")
    try:
        shutil.copytree(os.path.join(REF, "tests"), os.path.join(scratch, "tests"))
        env = dict(os.environ)
        env["PYTHONPATH"] = os.pathsep.join([os.path.join(HERE, "shims"), REF, scratch])
        cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider",
               "--rootdir", scratch, os.path.join(scratch, "tests"),
               "-k", "not benchmark"] + sys.argv[1:]
        return subprocess.call(cmd, cwd=scratch, env=env)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
