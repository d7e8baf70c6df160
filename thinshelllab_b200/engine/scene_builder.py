"""Scene construction on the host (cold path, numpy): what `Scene.__init__(); Scene.init_all(); Scene.reset()` leave in the reference's
fields, built from the mesh assets and the scenes' own formulas -- positions, masses, frozen flags, surface triangles, tetrahedral
cells with their rest matrices, the gripper frame -- as the `state` mapping the B200 scene classes consume.

Mirrors (cited per function): Cloth.__init__ / init_pos_offset* (code/engine/model_fold_offset.py:11-32, 826-868), the tactile
Elastic (code/engine/model_elastic_tactile.py:13-79, 215-229, 260-325), the box Elastic (code/engine/model_elastic_offset.py via
meshes.box_body), gripper.init_kernel (code/engine/gripper_single.py:51-78), BaseScene.__init__ / init_property / init_faces
(code/engine/BaseScene.py:30-135, 330-383) and the scenes' init_objects / init / set_frozen_kernel
(code/task_scene/Scene_folding.py:37-127, Scene_forming.py).  Rest angles (Cloth.init_ref_angle) are computed by the library
(tsl_cloth_update_ref_angle) when the scene object binds the state."""
import numpy as np

from ..meshes import box_body
from . import readfile


class TactileBody:
    """model_elastic_tactile.Elastic(dt, offset, ratio): the pad / ball mesh from TetGen files, scaled by `ratio`"""

    def __init__(self, ratio, name="tactile", E=300000.0, nu=0.2, density=2000.0):
        self.ratio, self.density = float(ratio), float(density)
        self.mu = E / (2 * (1 + nu))
        self.lam = E * nu / ((1 + nu) * (1 - 2 * nu))
        self.alpha = 1 + self.mu / self.lam
        n, ox = readfile.read_node(f"../data/{name}.node")
        _, cells = readfile.read_ele(f"../data/{name}.ele")
        _, f2v = readfile.read_smesh(f"../data/{name}.face")
        self.F_ox = np.asarray(ox, np.float64)
        self.tets = np.asarray(cells, np.int32)
        self.f2v_file = np.asarray(f2v, np.int32)
        self.n_verts, self.n_cells, self.n_surfaces = n, self.tets.shape[0], self.f2v_file.shape[0]
        self.is_surface = np.zeros(n, bool)                    # count() :302-311
        self.is_surface[self.f2v_file.reshape(-1)] = True
        r = np.linalg.norm(self.F_ox, axis=1)
        self.bottom = (self.F_ox[:, 2] < 0.001) & self.is_surface                    # is_bottom :253-254
        self.inner_circle = (r < 0.0076) & self.is_surface                            # is_inner_circle :257-258
        self.surf = (r > 0.0148) & self.is_surface                                    # is_surf :261-262

    def init(self, offset, flip):
        """Elastic.init -> init_pos + init_surface_indices (:215-229, 264-291): world positions, B = Ds^-1, W = |det Ds| / 6, lumped
        masses, surface triangles oriented outwards (inwards on the inner circle)"""
        off = np.asarray(offset, np.float64)
        x = self.ratio * self.F_ox
        if flip:
            x = -x
        self.F_x = x + off
        Ds = np.stack([self.F_x[self.tets[:, i]] - self.F_x[self.tets[:, 3]] for i in range(3)], -1)
        self.F_B = np.linalg.inv(Ds)
        self.F_W = np.abs(np.linalg.det(Ds)) / 6
        self.F_m = np.zeros(self.n_verts)
        np.add.at(self.F_m, self.tets.reshape(-1), np.repeat(self.F_W / 4 * self.density, 4))
        f = self.f2v_file.copy()
        p1, p2, p3 = self.F_x[f[:, 0]], self.F_x[f[:, 1]], self.F_x[f[:, 2]]
        n = np.cross(p2 - p1, p3 - p1)
        n /= np.linalg.norm(n, axis=1, keepdims=True)
        inner = off + np.array([0.0, 0.0, (-0.002 if flip else 0.002) * self.ratio])
        towards = np.einsum("ij,ij->i", n, inner[None] - p1) > 0
        all_inner = self.inner_circle[f].all(1)
        swap = towards != all_inner                            # swap when (towards and not inner) or (not towards and inner)
        f[swap, 1], f[swap, 2] = self.f2v_file[swap, 2], self.f2v_file[swap, 1]
        self.f2v = f
        return self


def cloth_positions_flat(N, M, dx, offset):
    """Cloth.init_pos_offset (:826-831)"""
    i, j = np.meshgrid(np.arange(N + 1), np.arange(M + 1), indexing="ij")
    return np.stack([i * dx + offset[0], j * dx + offset[1], np.full(i.shape, offset[2], np.float64)], -1).reshape(-1, 3)


def cloth_positions_fold(N, M, dx, offset, half_curv_num):
    """Cloth.init_pos_offset_fold (:841-861): the strip folded back over itself through a half circle (3.1415 as the reference has it)"""
    ox, oy, oz = offset
    r = dx
    if half_curv_num != 2:
        r = dx * (half_curv_num * 2 - 1) / 3.1415
    L, R = 7 - half_curv_num + 1, 7 + half_curv_num
    pos = np.zeros((N + 1, M + 1, 3))
    for i in range(N + 1):
        for j in range(M + 1):
            if i <= L:
                pos[i, j] = [(15 - i) * dx + ox, j * dx + oy, oz + 2 * r]
            if L + 1 <= i <= R - 1:
                x = (15 - L) * dx
                ang = (i - L) / (half_curv_num * 2 - 1) * 3.1415
                pos[i, j] = [x - r * np.sin(ang) + ox, j * dx + oy, oz + r * (1 + np.cos(ang))]
            if i >= R:
                pos[i, j] = [i * dx + ox, j * dx + oy, oz]
    return pos.reshape(-1, 3)


def pad_scene_state(*, cloth_size, cloth_N, cloth_M, cloth_pos, dt=5e-3, k_contact, eps_contact=0.0004, eps_v=0.01, max_n_constraints=10000,
                    rho=40.0, Kb=100.0, k_angle=3.14, mu=1.0, gravity=(0.0, 0.0, 0.0), table=(0.07, 9, 9, 2, (-0.035, -0.035, -0.00875)),
                    pad_ratio=0.015 / 0.03, pad_pos=None, pinned_last_row=True, init_ref_angle=True, cloth_topology=None):
    """state of a one-cloth / table / one-pad scene (Scene_folding, Scene_forming): BaseScene.__init__ + init_objects + init +
    init_property + set_frozen_kernel.  cloth_topology: f2v [NF,3] of the cloth (the library's, ShellEngine.cloth_topology)."""
    dx = cloth_size / cloth_N
    NVc = (cloth_N + 1) * (cloth_M + 1)
    tpos, ttets, tfaces, tmass = box_body(table[0], table[1], table[2], table[3], table[4])
    to, tn = NVc, tpos.shape[0]
    pad = TactileBody(pad_ratio).init(pad_pos, True)
    po, pn = to + tn, pad.n_verts
    pos0 = np.concatenate([cloth_pos, tpos, pad.F_x])
    mass = np.concatenate([np.full(NVc, rho * dx * dx), tmass, pad.F_m])
    frozen = np.zeros((pos0.shape[0], 3), np.int32)
    frozen[to:to + tn] = 1
    frozen[po:po + pn][pad.bottom | pad.inner_circle] = 1
    if pinned_last_row:
        frozen[cloth_N * (cloth_M + 1):NVc] = 1
    st = dict(dt=dt, k_contact=float(k_contact), eps_contact=eps_contact, eps_v=eps_v, mu=mu, Kb=Kb, k_angle=k_angle, cloth_N=cloth_N, cloth_M=cloth_M,
              cloth_dx=dx, cloth_mass=rho * dx * dx, cloth_size=cloth_size, max_n_constraints=max_n_constraints,
              pos0=pos0, vel0=np.zeros_like(pos0), mass=mass, frozen=frozen.reshape(-1), border_flag=np.zeros(pos0.shape[0], np.int32),
              gravity=np.asarray(gravity, np.float64), ref_angle0=np.zeros((2 * cloth_N * cloth_M, 3)), init_ref_angle=bool(init_ref_angle),
              table_tets=ttets, table_offset=to, table_nverts=tn, table_mu=5e5 / 2, table_lam=0.0,                    # box Elastic: E = 5e5, nu = 0 (model_elastic_offset.py:13-15)
              table_gravity=np.asarray(gravity, np.float64),
              pad_tets=pad.tets, pad_offset=po, pad_nverts=pn, pad_F_ox=pad.F_ox, pad_ratio=pad.ratio, pad_f2v=pad.f2v,
              pad_is_surface=pad.is_surface.astype(np.int32), pad_F_B=pad.F_B, pad_F_W=pad.F_W, pad_mu=pad.mu, pad_lam=pad.lam, pad_alpha=pad.alpha,
              pad_gravity=np.zeros(3),
              gripper_pos0=np.asarray(pad_pos, np.float64)[None], gripper_rot0=np.array([[1.0, 0.0, 0.0, 0.0]]),
              gripper_F_x=(pad.F_x - np.asarray(pad_pos, np.float64))[None],                     # gripper.init_kernel :55-60
              gripper_bound_idx=np.nonzero(pad.bottom | pad.inner_circle)[0].astype(np.int32),    # :62-68
              gripper_surface_idx=np.nonzero(pad.surf & ~(pad.bottom | pad.inner_circle))[0].astype(np.int32),
              _table_faces=tfaces + to, _pad_faces=pad.f2v + po)
    return st


def folding_state(cloth_size=0.06, forming=False, Kb=100.0):
    """Scene_folding (cloth 15 x 3, half_curve_num 2, k_contact 10000, k_angle 0.5) / Scene_forming (15 x 7, 3, 20000, 3.14)"""
    N, M = 15, (7 if forming else 3)
    hc = 3 if forming else 2
    dx = cloth_size / N
    cpos = cloth_positions_fold(N, M, dx, (-0.07, -0.02, 0.00035) if forming else (-0.07, -0.01, 0.0004), hc)
    r = dx * (hc * 2 - 1) / 3.1415
    x = -0.07 + (7 + hc) / 16 * 0.1 - r * 0.86 + (0.01 if forming else 0.005)
    pad_pos = (x, 0.0, 2 * r + (0.00785 if forming else 0.0079))
    return pad_scene_state(cloth_size=cloth_size, cloth_N=N, cloth_M=M, cloth_pos=cpos, k_contact=20000.0 if forming else 10000.0, Kb=Kb,
                           k_angle=3.14 if forming else 0.5, pad_pos=pad_pos)


# ------------------------------------------------------------------------------------------------ general one-cloth / many-bodies scenes
def multi_body_state(*, cloth_N, cloth_M, cloth_size, cloth_pos, elastics, pad_poses, dt=5e-3, k_contact, eps_contact=0.0004, eps_v=0.01,
                     max_n_constraints=10000, rho=40.0, Kb=100.0, k_angle=3.14, mu=1.0, cloth_gravity=(0.0, 0.0, 0.0), pinned_vertices=(),
                     init_ref_angle=False, mu_per_elastic=None, pad_part=None, pairs=None):
    """state of a scene with one cloth and a list of elastic bodies (BaseScene.__init__ + init_objects + init + init_property +
    set_frozen_kernel of Scene_lifting / Scene_pick).  elastics: list of dicts
        dict(kind="box", pos, tets, faces, mass, mu, lam, gravity, frozen=True/False)
        dict(kind="tactile", body=TactileBody (already .init()-ed), gravity)            -- driven by gripper part = its rank among the pads
        dict(kind="mesh", rest, pos, tets, faces, density, mu, lam, gravity)             -- a TetGen body (the ball), free
    pad_poses [n_parts][3]: gripper.init positions; pad_part[k] = gripper part driving pad k (default: pad k on part k; the two-finger
    gripper of gripper_tactile.py has pads 2 j and 2 j + 1 on part j).
    pairs: the scene's contact_analysis as a list of (surface body, vertex body, mu) in the reference's order (bodies: cloths first, then
    elastics); mu is a number or "elastic" / "cloth" (follows mu_cloth_elastic / mu_cloth_cloth), optionally ("elastic", factor).
    Default: every cloth against every elastic body both ways with mu_cloth_elastic (or mu_per_elastic[j])."""
    dx = cloth_size / cloth_N
    NVc = (cloth_N + 1) * (cloth_M + 1)
    # several cloths (Scene_card, Scene_sliding): cloth_pos is a list of [NVc][3] arrays, the cloths share N, M and material
    cloth_list = [np.asarray(cloth_pos, np.float64)] if np.ndim(cloth_pos[0][0]) == 0 else [np.asarray(p, np.float64) for p in cloth_pos]
    n_cloths = len(cloth_list)
    pos, mass, frozen = list(cloth_list), [np.full(NVc, rho * dx * dx)] * n_cloths, [np.zeros((NVc, 3), np.int32) for _ in range(n_cloths)]
    for v in pinned_vertices:
        frozen[0][v] = 1
    off = NVc * n_cloths
    els, faces = [], []
    for el in elastics:
        if el["kind"] == "mesh":
            p, tets, f = el["pos"], np.asarray(el["tets"], np.int32), el["faces"]
            B, W = tet_rest_np(el["rest"], tets)
            m = np.zeros(p.shape[0])
            np.add.at(m, tets.reshape(-1), np.repeat(W / 4 * el["density"], 4))          # Elastic.init_pos (:240-245)
            fz = np.zeros((p.shape[0], 3), np.int32)
            rec = dict(kind=0, offset=off, nverts=p.shape[0], tets=tets, F_B=B, F_W=W, mu=el["mu"], lam=el["lam"], alpha=0.0,
                       gravity=np.asarray(el["gravity"], np.float64), rest=p.copy())
        elif el["kind"] == "box":
            p, tets, f, m = el["pos"], el["tets"], el["faces"], el["mass"]
            B, W = tet_rest_np(p, tets)
            fz = np.full((p.shape[0], 3), 1 if el.get("frozen", False) else 0, np.int32)
            rec = dict(kind=0, offset=off, nverts=p.shape[0], tets=np.asarray(tets, np.int32), F_B=B, F_W=W, mu=el["mu"], lam=el["lam"], alpha=0.0,
                       gravity=np.asarray(el["gravity"], np.float64), rest=p.copy())
        else:
            b = el["body"]
            p, f, m = b.F_x, b.f2v, b.F_m
            fz = np.zeros((b.n_verts, 3), np.int32)
            fz[b.bottom | b.inner_circle] = 1
            rec = dict(kind=1, offset=off, nverts=b.n_verts, tets=b.tets, F_B=b.F_B, F_W=b.F_W, mu=b.mu, lam=b.lam, alpha=b.alpha,
                       gravity=np.asarray(el["gravity"], np.float64), bound_idx=np.nonzero(b.bottom | b.inner_circle)[0].astype(np.int32))
        pos.append(p); mass.append(m); frozen.append(fz); faces.append(np.asarray(f, np.int32) + off)
        els.append(rec)
        off += rec["nverts"]
    pads = [r for r in els if r["kind"] == 1]
    pos0 = np.concatenate(pos)
    pad_part = list(range(len(pads))) if pad_part is None else [int(v) for v in pad_part]
    n_parts = (max(pad_part) + 1) if pads else 0
    st = dict(dt=dt, k_contact=float(k_contact), eps_contact=eps_contact, eps_v=eps_v, mu=mu, Kb=Kb, k_angle=k_angle, cloth_N=cloth_N, cloth_M=cloth_M,
              cloth_dx=dx, cloth_mass=rho * dx * dx, cloth_size=cloth_size, max_n_constraints=max_n_constraints, pos0=pos0, vel0=np.zeros_like(pos0),
              mass=np.concatenate(mass), frozen=np.concatenate(frozen).reshape(-1), border_flag=np.zeros(pos0.shape[0], np.int32),
              gravity=np.asarray(cloth_gravity, np.float64), n_cloths=n_cloths,
              ref_angle0=np.zeros((2 * cloth_N * cloth_M, 3)) if n_cloths == 1 else np.zeros((n_cloths, 2 * cloth_N * cloth_M, 3)),
              init_ref_angle=bool(init_ref_angle),
              elastics=els, elastic_faces=faces, mu_per_elastic=mu_per_elastic,
              gripper_pos0=np.asarray(pad_poses, np.float64), gripper_rot0=np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (n_parts, 1)),
              pad_part=np.asarray(pad_part, np.int32),
              gripper_F_x=np.stack([pos0[r["offset"]:r["offset"] + r["nverts"]] - np.asarray(pad_poses[pad_part[k]], np.float64)
                                    for k, r in enumerate(pads)]),
              gripper_bound_idx=pads[0]["bound_idx"] if pads else np.zeros(0, np.int32))
    if pairs is not None:
        st["pairs"] = pairs
    return st


def tet_rest_np(rest, tets):
    """Elastic.init_pos: B = Ds^-1, W = |det Ds| / 6 of the rest positions"""
    tets = np.asarray(tets)
    D = np.stack([rest[tets[:, i]] - rest[tets[:, 3]] for i in range(3)], -1)
    return np.linalg.inv(D), np.abs(np.linalg.det(D)) / 6


def lifting_state(cloth_size=0.06, Kb=100.0):
    """Scene_lifting (code/task_scene/Scene_lifting.py:33-101, 136-151): a flat 15 x 15 cloth carrying a small heavy neo-Hookean box (free,
    under gravity), one pad above and two below, each on its own gripper part; cloth and pads carry no gravity; k_contact 500"""
    N = 15
    dx = cloth_size / N
    cpos = cloth_positions_flat(N, N, dx, (-0.03, -0.03, 0.0))
    bpos, btets, bfaces, bmass = box_body(0.007, 5, 5, 5, (-0.025, -0.005, 0.0003), density=20000.0)
    poses = [(0.01, 0.0, 0.0079), (0.0, -0.015, -0.0079), (0.0, 0.015, -0.0079)]
    flips = [True, False, False]
    els = [dict(kind="box", pos=bpos, tets=btets, faces=bfaces, mass=bmass, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, -9.8), frozen=False)]
    for p, fl in zip(poses, flips):
        els.append(dict(kind="tactile", body=TactileBody(0.015 / 0.03).init(p, fl), gravity=(0.0, 0.0, 0.0)))
    return multi_body_state(cloth_N=N, cloth_M=N, cloth_size=cloth_size, cloth_pos=cpos, elastics=els, pad_poses=poses, k_contact=500.0, Kb=Kb,
                            k_angle=3.14)


def pick_state(cloth_size=0.06, Kb=100.0):
    """Scene_pick (code/task_scene/Scene_pick.py:30-100): a flat 16 x 16 cloth over an ARCHED frozen table (friction 0.1 against it), two
    pads above on two gripper parts; gravity on the cloth; k_angle 0.5 (the task creases the cloth: training/trajopt_pick_fold.py)"""
    N = 16
    dx = cloth_size / N
    cpos = cloth_positions_flat(N, N, dx, (-0.03, -0.03, 0.0004))
    tpos, ttets, tfaces, tmass = box_body(0.06, 16, 16, 2, (-0.03, -0.03, -0.008), arch=0.004)
    poses = [(-0.025, 0.0, 0.0079), (0.025, 0.0, 0.0079)]
    els = [dict(kind="box", pos=tpos, tets=ttets, faces=tfaces, mass=tmass, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, -9.8), frozen=True)]
    for p in poses:
        els.append(dict(kind="tactile", body=TactileBody(0.015 / 0.03).init(p, True), gravity=(0.0, 0.0, 0.0)))
    return multi_body_state(cloth_N=N, cloth_M=N, cloth_size=cloth_size, cloth_pos=cpos, elastics=els, pad_poses=poses, k_contact=10000.0, Kb=Kb,
                            k_angle=0.5, cloth_gravity=(0.0, 0.0, -9.8), mu_per_elastic=[0.1, None, None])


def ball_body(pos, inner=None):
    """Elastic(load=True).init (code/engine/model_elastic_offset.py:40-49, 379-405): the TetGen ball of data/ball.*, NOT rescaled,
    translated to `pos`; surface triangles re-oriented to point away from `pos` (init_normal)"""
    from . import readfile
    _, verts = readfile.read_node("../data/ball.node")
    _, tets = readfile.read_ele("../data/ball.ele")
    _, faces = readfile.read_smesh("../data/ball.face")
    rest = np.asarray(verts, np.float64)
    x = rest + np.asarray(pos, np.float64)
    f = np.asarray(faces, np.int32).copy()
    p1, p2, p3 = x[f[:, 0]], x[f[:, 1]], x[f[:, 2]]
    n = np.cross(p2 - p1, p3 - p1)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    flip = np.einsum("ij,ij->i", n, np.asarray(pos if inner is None else inner, np.float64) - p1) > 0
    f[flip, 1], f[flip, 2] = f[flip, 2].copy(), f[flip, 1].copy()
    return rest, x, np.asarray(tets, np.int32), f


def balancing_state(cloth_size=0.06, Kb=100.0):
    """Scene_balancing (code/task_scene/Scene_balancing.py:27-96): a 15 x 7 cloth strip with the TetGen ball (free, density 10000, under
    gravity) resting on it, held at both ends between an upper and a lower tactile pad of a two-finger gripper (gripper_tactile.py): two
    parts, four pads; friction 0.2 against the ball; eps_contact 0.41 mm"""
    N, M = 15, 7
    dx = cloth_size / N
    cpos = cloth_positions_flat(N, M, dx, (-0.03, -0.015, 0.0))
    rest, bpos, btets, bfaces = ball_body((0.0, 0.0, 0.0039))
    parts = [(0.023, 0.0, 0.0), (-0.023, 0.0, 0.0)]
    pads = [((0.023, 0.0, 0.0079), True), ((0.023, 0.0, -0.0079), False), ((-0.023, 0.0, 0.0079), True), ((-0.023, 0.0, -0.0079), False)]
    els = [dict(kind="mesh", rest=rest, pos=bpos, tets=btets, faces=bfaces, density=10000.0, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, -9.8))]
    for p, fl in pads:
        els.append(dict(kind="tactile", body=TactileBody(0.015 / 0.03).init(p, fl), gravity=(0.0, 0.0, 0.0)))
    return multi_body_state(cloth_N=N, cloth_M=M, cloth_size=cloth_size, cloth_pos=cpos, elastics=els, pad_poses=parts, pad_part=[0, 0, 1, 1],
                            k_contact=10000.0, eps_contact=0.00041, Kb=Kb, k_angle=3.14, cloth_gravity=(0.0, 0.0, -9.8),
                            mu_per_elastic=[0.2, None, None, None, None])


def interact_state(cloth_size=0.06, Kb=100.0, dense=10000.0):
    """Scene_interact (code/task_scene/Scene_interact.py:28-104): a 15 x 15 cloth half on a frozen table, its free end between the two pads
    of ONE two-finger gripper part (which closes by 0.6 mm per frame over the first frames: Scene.action), and a free 6 x 6 x 4 box lying
    on the cloth; friction 0.2 against table and box, the box also touches the table (0.1); k_contact 30000"""
    N = 15
    dx = cloth_size / N
    cpos = cloth_positions_flat(N, N, dx, (-0.045, -0.03, 0.0004))
    tpos, ttets, tfaces, tmass = box_body(0.06, 16, 16, 2, (-0.03, -0.03, -0.004))
    bpos, btets, bfaces, bmass = box_body(0.012, 6, 6, 4, (0.001, -0.006, 0.0008), density=dense)
    els = [dict(kind="box", pos=tpos, tets=ttets, faces=tfaces, mass=tmass, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, -9.8), frozen=True),
           dict(kind="tactile", body=TactileBody(0.015 / 0.03).init((-0.04, 0.0, 0.0083), True), gravity=(0.0, 0.0, 0.0)),
           dict(kind="tactile", body=TactileBody(0.015 / 0.03).init((-0.04, 0.0, -0.0075), False), gravity=(0.0, 0.0, 0.0)),
           dict(kind="box", pos=bpos, tets=btets, faces=bfaces, mass=bmass, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, -9.8), frozen=False)]
    st = multi_body_state(cloth_N=N, cloth_M=N, cloth_size=cloth_size, cloth_pos=cpos, elastics=els, pad_poses=[(-0.04, 0.0, 0.0004)],
                          pad_part=[0, 0], k_contact=30000.0, Kb=Kb, k_angle=3.14, cloth_gravity=(0.0, 0.0, -9.8),
                          mu_per_elastic=[0.2, None, None, 0.2])
    st["extra_pairs"] = [(1, 4, 0.1), (4, 1, 0.1)]       # (surface of body a, vertices of body b, mu): table <-> box (contact_analysis :99-102)
    return st


def card_state(cloth_size=0.06, Kb=100.0):
    """Scene_card (code/task_scene/Scene_card.py:31-128): three stacked 12 x 8 cards on a frozen table, two pads at the ends turned by
    +-90 degrees about y (their faces look along x) and one pad above; the pads only act on cloth vertices (one-way pairs, 10 x the
    friction for cards 1 and 2), neighbouring cards touch each other with friction 0.1; k_contact 20000; damping 0.95"""
    N, M = 12, 8
    dx = cloth_size / N
    cpos = [cloth_positions_flat(N, M, dx, (-0.02, -0.02, z)) for z in (0.01, 0.0104, 0.0108)]
    tpos, ttets, tfaces, tmass = box_body(0.07, 9, 9, 2, (-0.025, -0.025, -0.00875))
    poses = [(-0.0285, 0.0, 0.01), (0.0485, 0.0, 0.01), (0.01, 0.0, 0.0185)]
    els = [dict(kind="box", pos=tpos, tets=ttets, faces=tfaces, mass=tmass, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, 0.0), frozen=True)]
    for p, fl in zip(poses, (False, False, True)):
        els.append(dict(kind="tactile", body=TactileBody(0.015 / 0.03).init(p, fl), gravity=(0.0, 0.0, 0.0)))
    pairs = []
    for i in range(3):
        for j in range(3):
            if abs(i - j) == 1:
                pairs += [(i, j, 0.1), (j, i, 0.1)]
    for i in range(3):
        for j in range(4):
            pairs.append((3 + j, i, "elastic" if i == 0 else ("elastic", 10.0)))
    s2 = np.sqrt(2.0) * 0.5
    st = multi_body_state(cloth_N=N, cloth_M=M, cloth_size=cloth_size, cloth_pos=cpos, elastics=els, pad_poses=poses, k_contact=20000.0, Kb=Kb,
                          k_angle=3.14, pairs=pairs)
    st["damping"] = 0.95
    st["gripper_rot0"] = np.array([[s2, 0.0, s2, 0.0], [s2, 0.0, -s2, 0.0], [1.0, 0.0, 0.0, 0.0]])    # init :88-92, then update_all:
    from .gripper_single import quat_to_rotmat32                  # every pad vertex to pos + R F_x (R held in float32 by the reference)
    pads = [r for r in st["elastics"] if r["kind"] == 1]
    for k, r in enumerate(pads):
        R = quat_to_rotmat32(st["gripper_rot0"][k]).astype(np.float64)
        st["pos0"][r["offset"]:r["offset"] + r["nverts"]] = np.asarray(poses[k]) + st["gripper_F_x"][k] @ R.T
    return st


def sliding_state(cloth_size=0.06, Kb=100.0):
    """Scene_sliding (code/task_scene/Scene_sliding.py:23-100): three stacked 15 x 15 cloths on a frozen table (friction 0.4 against it),
    one pad above (E 5e5 / nu 0.2 as every tactile pad); neighbouring cloths touch with the identified coefficient mu_cloth_cloth"""
    N = 15
    dx = cloth_size / N
    cpos = [cloth_positions_flat(N, N, dx, (-0.03, -0.03, z)) for z in (0.0004, 0.0008, 0.0012)]
    tpos, ttets, tfaces, tmass = box_body(0.1, 16, 16, 2, (-0.05, -0.05, -0.00666))
    poses = [(0.0, 0.0, 0.0105)]
    els = [dict(kind="box", pos=tpos, tets=ttets, faces=tfaces, mass=tmass, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, 0.0), frozen=True),
           dict(kind="tactile", body=TactileBody(0.015 / 0.03).init(poses[0], True), gravity=(0.0, 0.0, 0.0))]
    pad = els[1]["body"]                                           # Scene.__init__ :26-32: E 5e5, nu 0.2 for this pad
    E, nu = 5e5, 0.2
    pad.mu, pad.lam = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    pad.alpha = 1 + pad.mu / pad.lam
    pairs = []
    for i in range(3):
        for j in range(3):
            if abs(i - j) == 1:
                pairs += [(i, j, "cloth"), (j, i, "cloth")]
    n_cc = len(pairs)
    for i in range(3):
        for j in range(2):
            mu = 0.4 if j == 0 else "elastic"
            pairs += [(i, 3 + j, mu), (3 + j, i, mu)]
    st = multi_body_state(cloth_N=N, cloth_M=N, cloth_size=cloth_size, cloth_pos=cpos, elastics=els, pad_poses=poses, k_contact=10000.0, Kb=Kb,
                          k_angle=3.14, pairs=pairs)
    st["n_cloth_cloth_pairs"] = n_cc
    return st
