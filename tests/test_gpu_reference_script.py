"""The reference's system-identification driver (code/training/trajopt_bouncing.py:43-121) on the B200 engine, through the
module names the script itself imports (thinshelllab_b200.compat), against the numbers the SAME script printed with the
reference's own engine (tests/golden/trajopt_bouncing_T3.json).  If the script file travelled with the snapshot
(baseline/_ref/, git-ignored) it is also executed unmodified."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "trajopt_bouncing_T3.json")))


def test_driver_call_sequence_matches_reference_run():
    from thinshelllab_b200 import compat
    compat.install()
    import taichi as ti
    ti.init(ti.cpu, default_fp=ti.f64, default_ip=ti.i32, fast_math=False)
    from thinshelllab.engine.analytic_grad_system import Grad
    from thinshelllab.engine.geometry import projection_query
    from thinshelllab.task_scene.Scene_bouncing import Scene
    T, lr = GOLD["args"]["tot_step"], GOLD["args"]["lr"]
    sys_ = Scene(cloth_size=0.06)
    sys_.cloths[0].Kb[None] = GOLD["args"]["Kb"]
    analy_grad = Grad(sys_, T, sys_.elastic_cnt - 1)
    sys_.init_all()
    analy_grad.init_mass(sys_)
    sys_.reset()
    sys_.mu_cloth_elastic[None] = GOLD["args"]["mu_cloth_elastic"]
    analy_grad.copy_pos(sys_, 0)
    for frame in range(1, T):
        sys_.time_step(projection_query, frame)
        analy_grad.copy_pos(sys_, frame)
    reward = sys_.compute_reward()
    analy_grad.get_loss_table(sys_)
    for j in range(T - 1, 0, -1):
        analy_grad.transfer_grad(j, sys_, projection_query)
    grad = analy_grad.grad_kb[None] * lr
    assert abs(reward - GOLD["total_reward"]) < 1e-8, (reward, GOLD["total_reward"])
    assert abs(grad - GOLD["now_grad"]) <= 1e-5 * abs(GOLD["now_grad"]), (grad, GOLD["now_grad"])


def test_unmodified_script_if_present(tmp_path):
    script = os.path.join(ROOT, "baseline", "_ref", "trajopt_bouncing.py")
    if not os.path.exists(script):
        pytest.skip("the reference script is not part of this repository (git-ignored copy absent)")
    env = dict(os.environ, TSL_WORKDIR=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), script, "--l", "0", "--r", "1", "--iter", "1",
                          "--tot_step", "3"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    reward = float(re.search(r"total_reward: \[(?:np\.float64\()?([-0-9.e]+)", out.stdout).group(1))
    grad = float(re.search(r"now grad ([-0-9.e]+)", out.stdout).group(1))
    assert abs(reward - GOLD["total_reward"]) < 1e-8
    assert abs(grad - GOLD["now_grad"]) <= 1e-5 * abs(GOLD["now_grad"])
