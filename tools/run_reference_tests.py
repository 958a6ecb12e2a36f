#!/usr/bin/env python
"""Run the REFERENCE's own test suite against the drop-in class (only where /root/reference exists).

A shim package named ``picovdb`` re-exports ``picovdb_b200`` so the reference's tests import this
repository's ``PicoVectorDB`` unchanged.  ``--engine host`` (default, no GPU needed) plugs in the
test-only host engine from tests/_host_engine.py, so this checks the HOST logic of the drop-in
class (API, record schema, storage files, quirks) against the reference's expectations;
``--engine cuda`` runs the same tests on the real CUDA engine (needs a B200 and the reference tree,
so only where both exist).  Nothing from the reference is copied: its tests are run in place.

    python tools/run_reference_tests.py [--engine host|cuda] [pytest args...]
"""
import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PICOVDB_REFERENCE", "/root/reference")

SHIM_INIT = '''\
import sys
sys.path.insert(0, {root!r})
sys.path.insert(0, {tests!r})
import picovdb_b200 as _impl
from picovdb_b200 import *  # noqa: F401,F403
from picovdb_b200 import PicoVectorDB, K_ID, K_VECTOR, K_METRICS, _HAS_FAISS  # noqa: F401
if {host!r}:
    from _host_engine import HostEngine
    PicoVectorDB._engine_factory = staticmethod(lambda dim, **kw: HostEngine(dim, **kw))
'''

SHIM_SUB = '''\
from picovdb_b200.db import *  # noqa: F401,F403
from picovdb_b200.db import _RWLock, _normalize, _to_c_f32, logger  # noqa: F401
import picovdb_b200.db as _db
import sys as _sys
_sys.modules[__name__] = _db   # tests that monkeypatch picovdb.pico_vdb.<name> must hit the real module
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="host", choices=["host", "cuda"])
    args, rest = ap.parse_known_args()
    if not os.path.isdir(os.path.join(REF, "tests")):
        print(f"reference tree not found at {REF}: nothing to run")
        return 0
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, ".pytest_cache") if os.path.isdir(os.path.join(ROOT, ".pytest_cache")) else None) as tmp:
        pkg = os.path.join(tmp, "picovdb")
        os.makedirs(pkg)
        with open(os.path.join(pkg, "__init__.py"), "w") as f:
            f.write(SHIM_INIT.format(root=ROOT, tests=os.path.join(ROOT, "tests"), host=args.engine == "host"))
        with open(os.path.join(pkg, "pico_vdb.py"), "w") as f:
            f.write(SHIM_SUB)
        env = dict(os.environ, PYTHONPATH=tmp + os.pathsep + os.environ.get("PYTHONPATH", ""), PYTHONDONTWRITEBYTECODE="1")
        cmd = [sys.executable, "-m", "pytest", os.path.join(REF, "tests"), "-q", "-p", "no:cacheprovider",
               "--rootdir", tmp, *rest]
        return subprocess.call(cmd, env=env, cwd=tmp)


if __name__ == "__main__":
    sys.exit(main())
