#!/usr/bin/env python
"""Runs one of the REFERENCE's own driver scripts (e.g. code/training/trajopt_bouncing.py), unmodified, on the B200 engine:

    python tools/run_reference_script.py /path/to/trajopt_bouncing.py --l 0 --r 1 --iter 2 --tot_step 5

`thinshelllab_b200.compat.install()` maps the module names the script imports onto this package; the script runs with a
scratch working directory laid out like the reference's (cwd = <work>/code, so that its "../imgs/..." outputs land in
<work>/imgs).  The script file itself is never part of this repository."""
import os
import runpy
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thinshelllab_b200 import compat  # noqa: E402

script = os.path.abspath(sys.argv[1])
compat.install()
work = os.environ.get("TSL_WORKDIR") or tempfile.mkdtemp(prefix="tsl_ref_script_")
os.makedirs(os.path.join(work, "code"), exist_ok=True)
os.makedirs(os.path.join(work, "imgs"), exist_ok=True)
os.chdir(os.path.join(work, "code"))
sys.argv = [script] + sys.argv[2:]
print(f"[run_reference_script] {script} in {work}", flush=True)
with open(script) as f:
    relative = any(ln.startswith("from ..") for ln in f)
# some of the reference's drivers (trajopt_lifting.py, ...) import their package relatively ("from ..agent import ..."): they only run as
# members of the package, so give them the package name they would have there; none of them guards on __name__
runpy.run_path(script, run_name="thinshelllab.training." + os.path.splitext(os.path.basename(script))[0] if relative else "__main__")
