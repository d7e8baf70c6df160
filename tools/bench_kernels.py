"""Times the assembly kernel classes in isolation (CUDA events inside libtsl, tsl_bench_kernel) on the landing sheet after two steps
(contacts present).  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split.  usage: python tools/bench_kernels.py [N]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thinshelllab_b200 import _lib  # noqa: E402
from thinshelllab_b200.synthetic import LANDING, sheet_scene  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 707
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
s = sheet_scene(N, **LANDING)
e = s.engine
for _ in range(2):
    st = s.time_step()
e.assemble(_lib.ASM_RESIDUAL | _lib.ASM_HESSIAN | _lib.ASM_NEWTON | _lib.ASM_SPD)
out = {"N": N, "n_tris": 2 * N * N, "contacts": st.n_contacts}
for name, what in (("energy", 2), ("residual", 3), ("hessian_pair_fast", 4), ("hessian_one_scatter", 7), ("mg_setup", 6), ("pcg_iteration", 0), ("spmv", 1)):
    e.bench_kernel(what, 3)
    out[name + "_us"] = 1e3 * e.bench_kernel(what, iters)
e.set_option(_lib.OPT_FAST_ASSEMBLY, 2)
for name, what in (("energy_tiles", 2), ("residual_tiles", 3)):
    e.bench_kernel(what, 3)
    out[name + "_us"] = 1e3 * e.bench_kernel(what, iters)
print(json.dumps(out))
