#!/usr/bin/env python
"""TEST INFRASTRUCTURE (oracle tier 1): runs one of the reference's driver scripts with the REFERENCE's own engine under the
Taichi emulation shim (build container only; needs /root/reference).  Only the renderer -- out of scope, and dependent on
GGUI / trimesh -- is replaced by an inert stand-in.  Used to produce the comparison log for
tools/run_reference_script.py (same script, B200 engine).

    python oracle/run_reference_script_emulated.py /root/reference/code/training/trajopt_bouncing.py --l 0 --r 1 --iter 1 --tot_step 3
"""
import os
import runpy
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "ti_emu"))
import hook  # noqa: E402

hook.install()


class _Renderer:
    def __init__(self, *a, **k):
        pass

    def set_save_dir(self, d):
        os.makedirs(d, exist_ok=True)

    def render(self, *a, **k):
        pass

    def end_rendering(self, *a, **k):
        pass


sys.modules["thinshelllab.engine.render_engine"] = types.ModuleType("thinshelllab.engine.render_engine")
sys.modules["thinshelllab.engine.render_engine"].Renderer = _Renderer
if "imageio" not in sys.modules:
    try:
        import imageio  # noqa: F401
    except Exception:
        sys.modules["imageio"] = types.ModuleType("imageio")
if "open3d" not in sys.modules:
    try:
        import open3d  # noqa: F401
    except Exception:
        # readfile.save_cloth_mesh (PLY export for visualisation, out of scope) is the only user
        class _Mesh:
            def compute_vertex_normals(self):
                pass
        o3d = types.ModuleType("open3d")
        o3d.geometry = types.SimpleNamespace(TriangleMesh=_Mesh)
        o3d.utility = types.SimpleNamespace(Vector3iVector=lambda a: a, Vector3dVector=lambda a: a)
        o3d.io = types.SimpleNamespace(write_triangle_mesh=lambda *a, **k: True)
        sys.modules["open3d"] = o3d
# no GPU in the build container: the reference's hard-wired "cuda:0" torch device (BaseScene.py:31, sparse_solver.py:11)
# is forced to "cpu"; the arithmetic (fp64 torch tensors handed to the SuperLU stand-in) is unchanged
import thinshelllab.engine.BaseScene as _bs  # noqa: E402
import thinshelllab.engine.sparse_solver as _ss  # noqa: E402

_bs_init, _ss_init = _bs.BaseScene.__init__, _ss.SparseMatrix.__init__


def _bs_cpu(self, *a, **k):
    k["device"] = "cpu"
    _bs_init(self, *a, **k)


def _ss_cpu(self, n, use_cg=False, device="cpu"):
    _ss_init(self, n, use_cg, "cpu")


_bs.BaseScene.__init__, _ss.SparseMatrix.__init__ = _bs_cpu, _ss_cpu
script = os.path.abspath(sys.argv[1])
work = tempfile.mkdtemp(prefix="tsl_ref_emu_")
os.makedirs(os.path.join(work, "imgs"), exist_ok=True)
# the reference opens ../data/* relative to code/: keep its own code directory as cwd, send ../imgs elsewhere is not
# possible (read-only tree), so run from a scratch "code" dir with a data symlink
os.makedirs(os.path.join(work, "code"), exist_ok=True)
os.symlink(os.path.join(os.path.dirname(hook.REF_CODE), "data"), os.path.join(work, "data"))
os.chdir(os.path.join(work, "code"))
sys.argv = [script] + sys.argv[2:]
print(f"[emulated reference] {script} in {work}", flush=True)
runpy.run_path(script, run_name="__main__")
