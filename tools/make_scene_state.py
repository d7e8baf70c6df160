#!/usr/bin/env python
"""Extracts the scene DESCRIPTION (what Scene.init_all(); Scene.reset() leaves in the reference's fields: positions, masses, frozen
flags, faces, cells and rest matrices, gripper frame) from a golden written by oracle/gen_goldens.py into a small state file the
host scene classes load by default:

    python tools/make_scene_state.py tests/golden/folding.npz thinshelllab_b200/data/scene_folding_cloth0p1.npz
"""
import sys

import numpy as np

KEYS = ["dt", "k_contact", "eps_contact", "eps_v", "mu", "Kb", "k_angle", "cloth_N", "cloth_M", "cloth_dx", "cloth_mass", "cloth_size", "pos0", "vel0",
        "mass", "frozen", "faces", "ref_angle0", "border_flag", "gravity", "table_tets", "table_offset", "table_nverts", "pad_tets", "pad_offset",
        "pad_nverts", "pad_F_B", "pad_F_W", "pad_mu", "pad_lam", "pad_alpha", "pad_gravity", "table_gravity", "table_mu", "table_lam", "gripper_pos0",
        "gripper_F_x", "gripper_bound_idx", "body_v", "body_f"]
g = np.load(sys.argv[1])
np.savez_compressed(sys.argv[2], **{k: g[k] for k in KEYS})
print("wrote", sys.argv[2])
