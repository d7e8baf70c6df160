"""Makes the reference's driver scripts (code/training/trajopt_*.py) importable against this engine WITHOUT editing them.

The scripts do `import taichi as ti; ti.init(...)`, `import imageio`, `import matplotlib.pyplot`, and
`from thinshelllab.<pkg>.<module> import <names>` ("Public interface").  `install()` registers, under
those names, (a) this package's mirrors of the reference modules and (b) inert stand-ins for modules that are out of scope
(renderer) or not installed in this image (taichi, imageio, matplotlib) -- a stand-in is only used when the real module
cannot be imported.  Nothing here computes physics: Scene / Grad forward to libtsl (CUDA)."""
import importlib
import os
import sys
import types


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _have(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


class _Renderer:
    """engine/render_engine.py:Renderer -- visualisation is out of scope: keeps the calls, makes
    the output directory the scripts write their plots into"""

    def __init__(self, scene_sys, env_name, option="Taichi", config_path=None):
        self.save_dir = None

    def set_save_dir(self, save_dir):
        self.save_dir = save_dir
        os.makedirs(save_dir, exist_ok=True)

    def render(self, *a, **k):
        pass

    def end_rendering(self, *a, **k):
        pass


def install():
    from .. import fields  # noqa: F401
    from ..agent import traj_opt_single
    from ..engine import BaseScene, analytic_grad_single, analytic_grad_system, geometry, gripper_single, gripper_tactile, readfile
    from ..optimizer import optim
    from ..task_scene import Scene_bouncing, Scene_folding, Scene_forming, Scene_lifting, Scene_pick, Scene_balancing, Scene_interact, Scene_card, Scene_sliding

    # ---- third-party modules the scripts import at top level
    if not _have("taichi"):
        ti = _module("taichi", cpu="cpu", gpu="gpu", cuda="cuda", f64="f64", f32="f32", i32="i32", i64="i64")
        ti.init = lambda *a, **k: None
        ti.min, ti.max = min, max
        ti.math = _module("taichi.math", sqrt=lambda x: x ** 0.5)
    if not _have("imageio"):
        _module("imageio", imread=lambda *a, **k: None, imwrite=lambda *a, **k: None, mimsave=lambda *a, **k: None)
    if not _have("matplotlib.pyplot"):
        mpl = _module("matplotlib")
        mpl.pyplot = _module("matplotlib.pyplot", plot=lambda *a, **k: None, savefig=lambda *a, **k: None, figure=lambda *a, **k: None,
                             clf=lambda *a, **k: None, close=lambda *a, **k: None, show=lambda *a, **k: None, legend=lambda *a, **k: None)
    # ---- the thinshelllab module tree, as the scripts name it
    pkg = _module("thinshelllab", __path__=[])
    for sub in ("task_scene", "engine", "agent", "optimizer"):
        setattr(pkg, sub, _module(f"thinshelllab.{sub}", __path__=[]))
    table = {
        "thinshelllab.task_scene.Scene_bouncing": Scene_bouncing,
        "thinshelllab.engine.geometry": geometry,
        "thinshelllab.engine.analytic_grad_system": analytic_grad_system,
        "thinshelllab.engine.analytic_grad_single": analytic_grad_single,
        "thinshelllab.engine.gripper_single": gripper_single,
        "thinshelllab.engine.gripper_tactile": gripper_tactile,
        "thinshelllab.engine.readfile": readfile,
        "thinshelllab.task_scene.Scene_folding": Scene_folding,
        "thinshelllab.task_scene.Scene_forming": Scene_forming,
        "thinshelllab.task_scene.Scene_lifting": Scene_lifting,
        "thinshelllab.task_scene.Scene_pick": Scene_pick,
        "thinshelllab.task_scene.Scene_balancing": Scene_balancing,
        "thinshelllab.task_scene.Scene_interact": Scene_interact,
        "thinshelllab.task_scene.Scene_card": Scene_card,
        "thinshelllab.task_scene.Scene_sliding": Scene_sliding,
        "thinshelllab.engine.BaseScene": BaseScene,
        "thinshelllab.agent.traj_opt_single": traj_opt_single,
        "thinshelllab.optimizer.optim": optim,
    }
    for name, mod in table.items():
        sys.modules[name] = mod
        parent, leaf = name.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, mod)
    re_mod = _module("thinshelllab.engine.render_engine", Renderer=_Renderer)
    sys.modules["thinshelllab.engine"].render_engine = re_mod
    la = _module("thinshelllab.engine.linalg")          # imported by the scripts, never called by them
    sys.modules["thinshelllab.engine"].linalg = la
