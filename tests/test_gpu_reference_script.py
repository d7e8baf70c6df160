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


# ---- trajectory optimisation driver (code/training/trajopt_folding.py:50-140): Scene_folding + analytic_grad_single + Adam_single
GOLD_FOLD_PATH = os.path.join(ROOT, "tests", "golden", "trajopt_folding_T3.json")


def _gold_fold():
    if not os.path.exists(GOLD_FOLD_PATH):
        pytest.skip("no golden log of the emulated reference run of trajopt_folding.py")
    return json.load(open(GOLD_FOLD_PATH))


def test_folding_driver_call_sequence_matches_reference_run():
    """the statements of trajopt_folding.py's optimisation loop, through the module names the script imports: the reward of the second
    iteration depends on the whole chain forward -> adjoint -> gripper gradient -> Adam step -> fix_action -> forward"""
    gold = _gold_fold()
    from thinshelllab_b200 import compat
    compat.install()
    from thinshelllab.agent.traj_opt_single import agent_trajopt
    from thinshelllab.engine.analytic_grad_single import Grad
    from thinshelllab.engine.geometry import projection_query
    from thinshelllab.optimizer.optim import Adam_single
    from thinshelllab.task_scene.Scene_folding import Scene
    T, lr, iters = gold["args"]["tot_step"], gold["args"]["lr"], gold["args"]["iter"]
    sys_ = Scene(cloth_size=0.1)
    sys_.cloths[0].Kb[None] = 400.0
    analy_grad = Grad(sys_, T, sys_.elastic_cnt - 1)
    adam = Adam_single((T, sys_.elastic_cnt - 1, 6), lr, 0.9, 0.9999, 1e-8)
    agent = agent_trajopt(T, sys_.elastic_cnt - 1, max_moving_dist=0.001)
    sys_.init_all()
    analy_grad.init_mass(sys_)
    sys_.reset()
    sys_.mu_cloth_elastic[None] = 5.0
    adam.reset()
    rewards = []
    for i in range(iters):
        analy_grad.copy_pos(sys_, 0)
        for frame in range(1, T):
            agent.get_action(frame)
            sys_.action(frame, agent.delta_pos, agent.delta_rot)
            sys_.time_step(projection_query, frame)
            analy_grad.copy_pos(sys_, frame)
        rewards.append(sys_.compute_reward(1.0, -1.0))
        analy_grad.get_loss_fold(sys_, 1.0, -1.0)
        for j in range(T - 1, 0, -1):
            analy_grad.transfer_grad(j, sys_, projection_query)
        sys_.reset()
        adam.step(agent.traj, analy_grad.gripper_grad)
        agent.fix_action(0.015)
        analy_grad.reset()
    for mine, ref in zip(rewards, gold["total_reward"]):
        assert abs(mine - ref) <= 1e-5 * max(abs(ref), 1e-3), (rewards, gold["total_reward"])
    if "final_traj" in gold:
        import numpy as np
        # (1e-9 m / rad absolute: the golden trajectory is exactly zero, see the golden's note)
        assert np.abs(agent.traj.to_numpy() - np.array(gold["final_traj"])).max() <= 1e-5 * np.abs(gold["final_traj"]).max() + 1e-9


def test_unmodified_folding_script_if_present(tmp_path):
    gold = _gold_fold()
    script = os.path.join(ROOT, "baseline", "_ref", "trajopt_folding.py")
    if not os.path.exists(script):
        pytest.skip("the reference script is not part of this repository (git-ignored copy absent)")
    env = dict(os.environ, TSL_WORKDIR=str(tmp_path))
    a = gold["args"]
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), script, "--l", "0", "--r", "1", "--iter", str(a["iter"]),
                          "--tot_step", str(a["tot_step"])], capture_output=True, text=True, env=env, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    last = [ln for ln in out.stdout.splitlines() if ln.startswith("total_reward:")][-1]
    rewards = [float(x) for x in re.findall(r"(-?\d+\.\d+(?:e[-+]?\d+)?)", last.replace("np.float64(", ""))]
    assert len(rewards) == len(gold["total_reward"])
    for mine, ref in zip(rewards, gold["total_reward"]):
        assert abs(mine - ref) <= 1e-5 * max(abs(ref), 1e-3), (rewards, gold["total_reward"])


def _run_script(name, tmp_path, extra):
    script = os.path.join(ROOT, "baseline", "_ref", name)
    if not os.path.exists(script):
        pytest.skip("the reference script is not part of this repository (git-ignored copy absent)")
    env = dict(os.environ, TSL_WORKDIR=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), script] + extra,
                         capture_output=True, text=True, env=env, timeout=1200)
    assert out.returncode == 0, out.stderr[-3000:]
    last = [ln for ln in out.stdout.splitlines() if ln.startswith("total_reward:")][-1]
    return [float(x) for x in re.findall(r"(-?\d+\.\d+(?:e[-+]?\d+)?)", last.replace("np.float64(", ""))]


def test_unmodified_forming_script_if_present(tmp_path):
    """the reference's own training/trajopt_forming.py, unedited, on this engine: Scene_forming built by the product, get_loss_push,
    trajectory adjoint, Adam step -- two optimisation iterations towards a target 1 mm below the initial strip"""
    import numpy as np
    from thinshelllab_b200.engine.scene_builder import folding_state
    st = folding_state(cloth_size=0.1, forming=True)
    target = st["pos0"][:16 * 8].copy()
    target[:, 2] -= 1e-3
    tpath = str(tmp_path / "target.npy")
    np.save(tpath, target)
    rewards = _run_script("trajopt_forming.py", tmp_path, ["--l", "0", "--r", "1", "--iter", "2", "--tot_step", "3", "--target_dir", tpath, "--lr", "2e-5"])
    assert len(rewards) == 2 and all(np.isfinite(rewards)) and rewards[0] < 0
    assert rewards[1] > rewards[0]                      # one (small: Adam's first step is lr x sign) step along the adjoint gradient moves the strip towards the target


def test_unmodified_lifting_script_if_present(tmp_path):
    """training/trajopt_lifting.py, unedited: Scene_lifting (free box, three pads on three gripper parts), get_loss_lift,
    apply_action_limit_grad"""
    import numpy as np
    rewards = _run_script("trajopt_lifting.py", tmp_path, ["--l", "0", "--r", "1", "--iter", "2", "--tot_step", "3"])
    assert len(rewards) == 2 and all(np.isfinite(rewards)) and rewards[0] < 0


def test_unmodified_pick_fold_script_if_present(tmp_path):
    """training/trajopt_pick_fold.py, unedited: Scene_pick (arched frozen table with friction 0.1, two pads on two gripper parts, bending
    plasticity), agent.init_traj_pick_fold, compute_reward_pick_fold, get_loss_pick_fold (rest-angle seeds), the adjoint of frames
    tot_step-1 .. 9 and apply_action_limit_grad"""
    import numpy as np
    rewards = _run_script("trajopt_pick_fold.py", tmp_path, ["--l", "0", "--r", "1", "--iter", "2", "--tot_step", "11", "--render", "1000"])
    assert len(rewards) == 2 and all(np.isfinite(rewards))


def test_unmodified_balancing_script_if_present(tmp_path):
    """training/trajopt_balancing.py, unedited: Scene_balancing (TetGen ball, two-finger grippers), first a `--save` run that writes the
    state directory with save_all (the data/balance_state the reference ships lacks rot.npy / state: load_all cannot read it there either),
    then the optimisation run that load_all()s it at the start of every rollout; get_loss_balance, apply_action_limit_grad"""
    import numpy as np
    state = str(tmp_path / "balance_state")
    common = ["--l", "0", "--r", "1", "--tot_step", "3", "--render", "1000", "--load_state", state]
    _run_script("trajopt_balancing.py", tmp_path, common + ["--iter", "1", "--save"])
    for name in ("F_x_upper.npy", "rot.npy", "half_gripper_dist.npy", "state", "border_flag.npy"):
        assert os.path.exists(os.path.join(state, name)), name
    rewards = _run_script("trajopt_balancing.py", tmp_path, common + ["--iter", "2"])
    assert len(rewards) == 2 and all(np.isfinite(rewards)) and rewards[0] < 0


def test_unmodified_interact_script_if_present(tmp_path):
    """training/trajopt_interact.py --sep, unedited: Scene_interact (closing two-finger gripper, free box, elastic-elastic contact),
    get_loss_interact, the adjoint of frames tot_step-1 .. 6, Adam with a discount"""
    import numpy as np
    rewards = _run_script("trajopt_interact.py", tmp_path, ["--l", "0", "--r", "1", "--iter", "2", "--tot_step", "8", "--sep", "--render", "1000"])
    assert len(rewards) == 2 and all(np.isfinite(rewards))


def test_unmodified_sliding_script_if_present(tmp_path):
    """training/trajopt_silding.py, unedited: Scene_sliding (three cloths, cloth-cloth contact), analytic_grad_system.Grad with
    count_friction_grad, agent.init_traj_slide, the update of mu_cloth_cloth"""
    import numpy as np
    rewards = _run_script("trajopt_silding.py", tmp_path, ["--l", "0", "--r", "1", "--iter", "2", "--tot_step", "8", "--mu", "0.5", "--lr", "1e-4"])
    assert len(rewards) == 2 and all(np.isfinite(rewards))


def test_unmodified_card_script_if_present(tmp_path):
    """training/trajopt_card.py, unedited: Scene_card (three cards), agent.init_traj_card, get_loss_card, the Kb update (the driver
    back-propagates frames above 50 only: with tot_step 6 this covers the rollout, the reward and the update arithmetic)"""
    import numpy as np
    rewards = _run_script("trajopt_card.py", tmp_path, ["--l", "0", "--r", "1", "--iter", "2", "--tot_step", "6"])
    assert len(rewards) == 2 and all(np.isfinite(rewards))
