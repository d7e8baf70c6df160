"""Scene construction (thinshelllab_b200/engine/scene_builder.py, engine/readfile.py: cold path, numpy) against the arrays the
REFERENCE produced for the same scenes (tests/golden/folding.npz, forming.npz: dumps of Scene.init_all(); Scene.reset() under the
Taichi emulation, oracle/gen_goldens.py:gen_folding): positions, masses, frozen flags, surface triangles with their orientation,
tetrahedral cells and rest matrices, the gripper frame."""
import os

import numpy as np
import pytest

from thinshelllab_b200.engine import readfile
from thinshelllab_b200.engine.scene_builder import TactileBody, folding_state


def test_tetgen_readers_match_the_reference_mesh(golden_dir):
    g = np.load(os.path.join(golden_dir, "folding.npz"))
    n, ox = readfile.read_node()
    m, cells = readfile.read_ele()
    k, faces = readfile.read_smesh()
    assert (n, m, k) == (276, 1365, 200)
    assert np.array_equal(np.asarray(ox), g["pad_F_ox"]) and np.array_equal(np.asarray(cells), g["pad_tets"])
    nb, _ = readfile.read_node("../data/ball.node")
    assert nb == 100                                      # data/ball.*: the ball of Scene_balancing (BASELINE configs[3])
    ball = TactileBody(1.0, name="ball").init((0.0, 0.0, 0.0), False)
    assert ball.F_W.min() > 0 and abs(ball.F_m.sum() - ball.F_W.sum() * ball.density) < 1e-12


@pytest.mark.parametrize("tag,forming", [("folding", False), ("forming", True)])
def test_scene_state_matches_the_reference(golden_dir, tag, forming):
    g = np.load(os.path.join(golden_dir, f"{tag}.npz"))
    st = folding_state(cloth_size=0.1, forming=forming)
    exact = ("pos0", "vel0", "frozen", "pad_F_B", "pad_f2v", "pad_is_surface", "gripper_pos0", "gripper_F_x", "gripper_bound_idx", "table_tets",
             "pad_tets", "border_flag", "gravity", "pad_gravity", "table_gravity", "cloth_dx", "pad_mu", "pad_lam", "pad_alpha", "table_mu", "table_lam")
    for k in exact:
        assert np.array_equal(np.asarray(st[k]), np.asarray(g[k])), k
    for k in ("mass", "pad_F_W", "cloth_mass"):
        assert np.abs(np.asarray(st[k]) - g[k]).max() <= 1e-14 * np.abs(g[k]).max(), k
    for k in ("k_contact", "eps_contact", "eps_v", "dt", "k_angle", "cloth_N", "cloth_M"):
        assert float(st[k]) == float(g[k]), k
    assert int(st["table_offset"]) == int(g["table_offset"]) and int(st["pad_offset"]) == int(g["pad_offset"])
    nfc = int(g["body_f"][int(g["n_cloths"]) - 1 if "n_cloths" in g.files else 0][1])
    faces = np.concatenate([g["faces"][:nfc], st["_table_faces"], st["_pad_faces"]])      # (cloth faces come from the library's mesher)
    assert np.array_equal(faces, g["faces"])


def test_other_cloth_sizes_build():
    for size in (0.06, 0.08):
        st = folding_state(cloth_size=size)
        assert st["pos0"].shape[0] == 64 + 162 + 276 and np.isfinite(st["pos0"]).all()
        assert abs(st["cloth_dx"] - size / 15) < 1e-18


@pytest.mark.parametrize("tag", ["lifting", "pick", "balancing", "interact", "card", "sliding"])
def test_multi_body_state_matches_the_reference(golden_dir, tag):
    """Scene_lifting (flat cloth + free heavy box + three pads on three gripper parts) and Scene_pick (cloth on an arched frozen table +
    two pads on two parts): engine/scene_builder.{lifting,pick}_state against the arrays Scene(); init_all(); reset() left in the
    reference (tests/golden/scene_state_{lifting,pick}.npz, oracle/gen_goldens.py scene_states)"""
    from thinshelllab_b200.engine import scene_builder
    g = np.load(os.path.join(golden_dir, f"scene_state_{tag}.npz"))
    st = getattr(scene_builder, f"{tag}_state")(0.06)
    for k in ("pos0", "vel0", "frozen", "gripper_pos0", "gripper_bound_idx", "border_flag"):
        assert np.array_equal(np.asarray(st[k]), g[k]), k
    if tag in ("balancing", "interact"):      # two-finger gripper: pads (upper, lower) per part, relative to the part's position
        assert np.array_equal(st["gripper_F_x"][0::2], g["gripper_F_x_upper"]) and np.array_equal(st["gripper_F_x"][1::2], g["gripper_F_x_lower"])
    else:
        assert np.array_equal(st["gripper_F_x"], g["gripper_F_x"])
    assert np.abs(st["mass"] - g["mass"]).max() <= 1e-14 * g["mass"].max()
    assert float(st["k_contact"]) == float(g["k_contact"]) and np.array_equal(st["gravity"], g["cloth_gravity"])
    assert len(st["elastics"]) == int(g["n_elastics"])
    for j, el in enumerate(st["elastics"]):
        assert el["offset"] == int(g[f"el{j}_offset"]) and el["kind"] == int(g[f"el{j}_tactile"])
        assert np.array_equal(el["tets"], g[f"el{j}_tets"]) and np.array_equal(el["gravity"], g[f"el{j}_gravity"])
        nc = el["tets"].shape[0]       # (the reference sizes F_B / F_W of the loaded ball by the box formula: 320 entries for 295 cells)
        assert np.abs(el["F_B"] - g[f"el{j}_F_B"][:nc]).max() <= 1e-13 * np.abs(g[f"el{j}_F_B"]).max()
        assert np.abs(el["F_W"] - g[f"el{j}_F_W"][:nc]).max() <= 1e-13 * g[f"el{j}_F_W"].max()
        assert float(el["mu"]) == float(g[f"el{j}_mu"]) and float(el["lam"]) == float(g[f"el{j}_lam"])
    nfc = int(g["body_f"][int(g["n_cloths"]) - 1 if "n_cloths" in g.files else 0][1])
    assert np.array_equal(np.concatenate([g["faces"][:nfc]] + st["elastic_faces"]), g["faces"])
