"""The module tree the reference's driver scripts import resolves against this package (no GPU needed to import), and the
host-side agent / optimiser mirrors keep the reference's semantics (code/agent/traj_opt_single.py:15-48,
code/optimizer/optim.py:37-81)."""
import sys

import numpy as np
import pytest
import torch


@pytest.fixture
def clean_modules():
    """compat.install() registers stand-ins in sys.modules: remove whatever it added, other tests import the real things"""
    before = set(sys.modules)
    yield
    for k in set(sys.modules) - before:
        if k.split(".")[0] in ("taichi", "imageio", "matplotlib", "thinshelllab"):
            del sys.modules[k]


def test_reference_import_names_resolve(clean_modules):
    from thinshelllab_b200 import compat
    compat.install()
    import taichi as ti
    ti.init(ti.cpu, default_fp=ti.f64, default_ip=ti.i32, fast_math=False)
    from thinshelllab.agent.traj_opt_single import agent_trajopt  # noqa: F401
    from thinshelllab.engine import linalg  # noqa: F401
    from thinshelllab.engine.analytic_grad_system import Grad  # noqa: F401
    from thinshelllab.engine.geometry import projection_query  # noqa: F401
    from thinshelllab.engine.render_engine import Renderer
    from thinshelllab.optimizer.optim import Adam_single  # noqa: F401
    from thinshelllab.task_scene.Scene_bouncing import Body, Scene  # noqa: F401
    import imageio  # noqa: F401
    import matplotlib.pyplot as plt
    plt.plot([0], [0])
    assert hasattr(Renderer, "set_save_dir") and hasattr(Renderer, "render") and hasattr(Renderer, "end_rendering")


def test_agent_and_adam_semantics():
    from thinshelllab_b200.agent.traj_opt_single import agent_trajopt
    from thinshelllab_b200.optimizer.optim import Adam_single
    a = agent_trajopt(4, 2, max_moving_dist=0.0005)
    a.traj[1, 0, 0] = 0.002                      # 4x too far: rescaled to 0.0005
    a.traj[2, 0, 0] = 0.0021                     # then measured from the rescaled frame 1
    a.fix_action(0.015)
    t = a.traj.to_numpy()
    assert abs(t[1, 0, 0] - 0.0005) < 1e-8 and abs(t[2, 0, 0] - 0.001) < 1e-8
    a.get_action(2)
    assert np.allclose(a.delta_pos.to_numpy()[0], [0.0005, 0, 0], atol=1e-8)
    # Adam: epsilon inside the square root, lr decays by `discount` every 10 steps
    p = torch.zeros((1, 1, 1), dtype=torch.float64)
    opt = Adam_single((1, 1, 1), 1e-2, 0.9, 0.999, 1e-8, discount=0.5)
    opt.step(p, torch.ones((1, 1, 1)))
    assert abs(float(p) + 1e-2 / np.sqrt(1 + 1e-8)) < 1e-12
    for _ in range(9):
        opt.step(p, torch.ones((1, 1, 1)))
    assert abs(opt.lr - 5e-3) < 1e-15


def test_gripper_pose_update_matches_reference(golden_dir):
    """gripper.step_simple + get_rotmat (code/engine/gripper_single.py:85-128) against the emulated reference rollout of
    Scene_folding: positions, quaternions (f64) and the float32 rotation matrices of frames 1 and 2"""
    import os
    import numpy as np
    from thinshelllab_b200.engine.gripper_single import pose_step, quat_to_rotmat32
    g = np.load(os.path.join(golden_dir, "folding.npz"))
    pos, rot = g["gripper_pos0"][0].copy(), g["gripper_rot0"][0].copy()
    traj = g["traj"]
    for frame in (1, 2):
        d = traj[frame, 0] - traj[frame - 1, 0]
        pos, rot = pose_step(pos, rot, d[:3], d[3:])
        assert np.abs(pos - g[f"f{frame}_gripper_pos"][0]).max() < 1e-15
        assert np.abs(rot - g[f"f{frame}_gripper_rot"][0]).max() < 1e-15
        R = quat_to_rotmat32(rot)
        assert R.dtype == np.float32 and np.array_equal(R, g[f"f{frame}_gripper_rotmat"][0])
