"""Case-directory fixtures (tests/golden/cases): the five b*.txt files and the
raw dump of on_fault_vars_input.nc, produced by the reference's own case
workflow (tools/gen_case_fixtures.py).  Large on-fault dumps are stored gzipped;
`materialize` gives a plain directory the host library can read."""
import gzip
import os
import shutil
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = os.path.join(ROOT, "tests", "golden", "cases")


def case_dir(name):
    d = name if os.path.isdir(name) else os.path.join(CASES, name)
    if not os.path.isdir(d):
        raise FileNotFoundError("no such case fixture: %s" % name)
    return d


def materialize(name, scratch=None):
    """Return a directory holding the plain input files of case `name`."""
    d = case_dir(name)
    gz = os.path.join(d, "on_fault_vars_input.bin.gz")
    if not os.path.exists(gz):
        return d
    out = scratch or os.path.join(tempfile.gettempdir(), "eqd_case_%s_%d" % (os.path.basename(d), os.getuid()))
    os.makedirs(out, exist_ok=True)
    for f in os.listdir(d):
        if f.endswith(".gz"):
            dst = os.path.join(out, f[:-3])
            if not os.path.exists(dst):
                tmp = dst + ".tmp%d" % os.getpid()
                with gzip.open(os.path.join(d, f), "rb") as a, open(tmp, "wb") as b:
                    shutil.copyfileobj(a, b)
                os.replace(tmp, dst)
        else:
            # several ranks may materialise the same case at once: only ever publish whole files
            src, dst = os.path.join(d, f), os.path.join(out, f)
            if os.path.exists(dst) and os.path.getsize(dst) == os.path.getsize(src):
                continue
            tmp = dst + ".tmp%d" % os.getpid()
            shutil.copy(src, tmp)
            os.replace(tmp, dst)
    return out
