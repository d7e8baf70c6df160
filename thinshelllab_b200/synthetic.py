"""Synthetic sheet scenes of BASELINE.json's configs: an N x N cloth over a frozen box table, Scene_bouncing physics.

Geometry is chosen so that the reference's contact query is meaningful at every size: cloth grid step `dx` (default
2 mm), table triangles no larger than the 3 mm contact grid cell, and the contact grid extent (`grid_n`, the
reference hard-codes 132 cells = +-0.1965 m, geometry.py:8-10) widened to cover the sheet.

`sheet_spec` is pure numpy (host, no GPU): the same description feeds the CUDA scene, the parity tests' oracle and
bench.py's CPU baseline arm."""
import numpy as np

from .meshes import box_body


def sheet_spec(N, dx=0.002, dt=5e-3, seed=0, z0=0.0003, bump=0.25, noise=0.01, k_contact=40000.0, mu=0.5):
    size = N * dx
    table_size = size + 0.02
    tn = int(np.ceil(table_size / 0.003)) + 1
    grid_n = max(132, 2 * int(np.ceil((0.5 * table_size + 0.01) / 0.003)) + 2)
    NV = (N + 1) ** 2
    tdx = table_size / (tn - 1)
    tpos, ttets, tfaces, tmass = box_body(table_size, tn, tn, 2, (-0.5 * table_size, -0.5 * table_size, -tdx))
    # deterministic start: flat sheet + smooth bump + small noise
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(N + 1), np.arange(N + 1), indexing="ij")
    cpos = np.stack([i * dx - 0.5 * size, j * dx - 0.5 * size, np.full(i.shape, z0, np.float64)], -1).reshape(-1, 3)
    cpos[:, 2] += (bump * dx * (1 + np.sin(2 * np.pi * i / 32.0) * np.cos(2 * np.pi * j / 32.0))).reshape(-1)
    cpos += rng.uniform(-noise * dx, noise * dx, (NV, 3))
    return dict(N=N, dx=dx, dt=dt, size=size, table_size=table_size, table_N=(tn, tn, 2), table_offset=(-0.5 * table_size, -0.5 * table_size, -tdx),
                table_pos=tpos, table_faces=tfaces, table_mass=tmass, cloth_pos=cpos, grid_n=grid_n, k_contact=k_contact, mu=mu,
                eps_contact=0.0004, eps_v=0.01, max_n_constraints=NV + 16, n_tris=2 * N * N, n_verts=NV + tpos.shape[0])


# The benchmark scenario (bench.py, BASELINE configs[1], [2], [4]) and the scenario of the BASELINE-size parity tests: the bumpy, noisy
# sheet released 0.3 mm above the table, i.e. inside the 0.4 mm contact gap -- a LANDING: contact from step 0, impact / rebound over
# steps 1-4, at rest afterwards.  FREE_FALL releases it outside the gap (0.6 mm): with 1 % in-plane noise and nothing to stop
# out-of-plane motion the free sheet wrinkles at element scale -- 302 Newton iterations on step 0 and an unconverged impact step at
# 1 M triangles (profiles/r2_bench707_freefall.json; the CPU oracle needs 224 on a 32 x 32 sheet) -- a solver stress test, not a benchmark.
LANDING = dict(z0=0.0003)
FREE_FALL = dict(z0=0.0006)


def sheet_scene(N, device="cuda:0", pinned_vertices=(), **kw):
    import torch

    from .task_scene.Scene_bouncing import Scene
    sp = sheet_spec(N, **kw)
    size = sp["size"]
    s = Scene(cloth_size=size, cloth_N=N, dt=sp["dt"], table_size=sp["table_size"], table_N=sp["table_N"], table_offset=sp["table_offset"],
              cloth_offset=(-0.5 * size, -0.5 * size, 0.0), reset_offset=(-0.5 * size, -0.5 * size, 0.0),
              k_contact=sp["k_contact"], max_n_constraints=sp["max_n_constraints"], grid_n=sp["grid_n"], pinned_vertices=pinned_vertices, device=device)
    s.mu_cloth_elastic[None] = sp["mu"]
    s.init_all()
    NV = (N + 1) ** 2
    s.engine.pos[:NV] = torch.from_numpy(sp["cloth_pos"]).to(s.engine.device)
    s.engine.prev_pos.copy_(s.engine.pos)
    s.engine.cloth_ref_angle[0].zero_()     # a flat sheet: no pre-creased rows (Scene_bouncing's init_ref_angle_bridge is scene specific)
    s.spec = sp
    return s


def pad_sheet_state(N, pad, dx=0.002, dt=5e-3, Kb=100.0, k_angle=3.14, k_contact=10000.0, mu=0.5, gap=2e-4):
    """BASELINE.json configs[3]-style scene as a state mapping for task_scene.Scene_folding.Scene: an N x N sheet resting on a frozen
    table with a tactile pad (volumetric, kinematic gripper) hovering `gap` above its centre, ready to be pressed into it.
    `pad` holds the pad's arrays with the key names of oracle/gen_goldens.py:gen_folding (pad_tets, pad_F_B, pad_F_W, pad_mu, pad_lam,
    pad_alpha, gripper_F_x, gripper_bound_idx, and faces / mass / frozen restricted to the pad through pad_offset / body_f)."""
    sp = sheet_spec(N, dx=dx, dt=dt, bump=0.0, noise=0.0, z0=0.0004, k_contact=k_contact, mu=mu)
    NVc, NFc = (N + 1) ** 2, 2 * N * N
    tpos, tfaces, tmass = sp["table_pos"], sp["table_faces"], sp["table_mass"]
    po, pn = int(pad["pad_offset"]), int(pad["pad_nverts"])
    Fx = np.asarray(pad["gripper_F_x"], np.float64)[0]
    if "faces" in pad:                          # a reference-made state (tests/golden/folding.npz)
        pf0, pf1 = (int(v) for v in np.asarray(pad["body_f"])[2])
        pfaces = np.asarray(pad["faces"])[pf0:pf1] - po
    else:                                       # engine/scene_builder.folding_state
        pfaces = np.asarray(pad["_pad_faces"]) - po
    gpos = np.array([[0.0, 0.0, 0.0004 + 0.0004 + gap - Fx[:, 2].min()]])
    ppos = gpos + Fx
    to, pno = NVc, NVc + tpos.shape[0]
    pos0 = np.concatenate([sp["cloth_pos"], tpos, ppos])
    mass = np.concatenate([np.full(NVc, 40.0 * dx * dx), tmass, np.asarray(pad["mass"])[po:po + pn]])
    frozen = np.zeros((pos0.shape[0], 3), np.int32)
    frozen[to:pno] = 1
    frozen[pno:] = np.asarray(pad["frozen"]).reshape(-1, 3)[po:po + pn]
    return dict(dt=dt, k_contact=k_contact, eps_contact=sp["eps_contact"], eps_v=sp["eps_v"], mu=mu, Kb=Kb, k_angle=k_angle, cloth_N=N, cloth_M=N,
                cloth_dx=dx, cloth_mass=40.0 * dx * dx, pos0=pos0, vel0=np.zeros_like(pos0), mass=mass, frozen=frozen.reshape(-1),
                _cloth_faces=None, _table_faces=tfaces + to, _pad_faces=pfaces + pno, ref_angle0=np.zeros((NFc, 3)),
                border_flag=np.zeros(pos0.shape[0], np.int32), gravity=np.array([0.0, 0.0, -9.8]),
                table_offset=to, table_nverts=tpos.shape[0], pad_tets=pad["pad_tets"], pad_offset=pno, pad_nverts=pn, pad_F_B=pad["pad_F_B"],
                pad_F_W=pad["pad_F_W"], pad_mu=pad["pad_mu"], pad_lam=pad["pad_lam"], pad_alpha=pad["pad_alpha"], pad_gravity=np.zeros(3),
                gripper_pos0=gpos, gripper_F_x=pad["gripper_F_x"], gripper_bound_idx=pad["gripper_bound_idx"],
                max_n_constraints=4 * NVc + 4096, grid_n=sp["grid_n"], n_tris=NFc)


def pad_sheet_scene(N, pad=None, device="cuda:0", **kw):
    from .task_scene.Scene_folding import Scene
    if pad is None:
        from .engine.scene_builder import folding_state
        pad = folding_state(cloth_size=0.1)
    st = pad_sheet_state(N, pad, **kw)
    return Scene(st, device=device, max_newton=200)


def config3_state(N, dx=0.002, dt=5e-3, Kb=100.0, k_contact=10000.0, mu=0.5, gap=2e-4, ball_xy=(0.03, 0.0), overhang=0.021):
    """BASELINE.json configs[3]: an N x N sheet + the ball of data/ball.* resting on it + two tactile pads of data/tactile.*
    gripping an edge (the Scene_balancing ingredients, scaled).  The sheet rests on a frozen table; its +x edge overhangs the table by
    `overhang` and sits between the upper and the lower pad of one two-finger gripper part (engine/gripper_tactile.py), each `gap` away
    from the sheet, ready to close and lift; the TetGen ball (free, density 10000, under gravity) lies on the sheet.  A state mapping for
    task_scene._multi_body.MultiBodyScene."""
    from .engine.scene_builder import TactileBody, ball_body, multi_body_state
    sp = sheet_spec(N, dx=dx, dt=dt, bump=0.0, noise=0.0, z0=0.0004, k_contact=k_contact, mu=mu)
    tn = sp["table_N"][0]
    tdx = sp["table_size"] / (tn - 1)
    cut = int(np.ceil((overhang + 0.01) / tdx))                          # table columns dropped on the +x side (the table is 1 cm wider than the sheet)
    tpos, ttets, tfaces, tmass = box_body(sp["table_size"], tn - cut, tn, 2, sp["table_offset"])
    assert tpos[:, 0].max() < 0.5 * N * dx - overhang + tdx
    zs = 0.0004                                                          # the sheet's plane
    px = 0.5 * N * dx - 0.008                                            # pad axis 8 mm inside the overhanging edge
    probe = TactileBody(0.015 / 0.03).init((0.0, 0.0, 0.0), True)
    reach = -probe.F_x[:, 2].min()                                       # distance from a pad's pose to its sensing tip
    pads = [TactileBody(0.015 / 0.03).init((px, 0.0, zs + 0.0004 + gap + reach), True),
            TactileBody(0.015 / 0.03).init((px, 0.0, zs - 0.0004 - gap - reach), False)]
    assert pads[1].F_x[:, 0].min() > tpos[:, 0].max()                    # the lower pad hangs beside the table, not inside it
    rest, bpos, btets, bfaces = ball_body((ball_xy[0], ball_xy[1], zs + 0.0039))
    els = [dict(kind="box", pos=tpos, tets=ttets, faces=tfaces, mass=tmass, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, -9.8), frozen=True),
           dict(kind="tactile", body=pads[0], gravity=(0.0, 0.0, 0.0)),
           dict(kind="tactile", body=pads[1], gravity=(0.0, 0.0, 0.0)),
           dict(kind="mesh", rest=rest, pos=bpos, tets=btets, faces=bfaces, density=10000.0, mu=5e5 / 2, lam=0.0, gravity=(0.0, 0.0, -9.8))]
    st = multi_body_state(cloth_N=N, cloth_M=N, cloth_size=N * dx, cloth_pos=sp["cloth_pos"], elastics=els, pad_poses=[(px, 0.0, zs)], pad_part=[0, 0],
                          dt=dt, k_contact=k_contact, Kb=Kb, k_angle=3.14, mu=mu, cloth_gravity=(0.0, 0.0, -9.8),
                          max_n_constraints=4 * (N + 1) ** 2 + 4096)
    st["grid_n"] = sp["grid_n"]
    st["n_tris"] = 2 * N * N
    return st


def config3_scene(N, device="cuda:0", **kw):
    from .task_scene._multi_body import MultiBodyScene

    class Scene(MultiBodyScene):
        max_newton = 200

        def __init__(self, st):
            self._build(st, device=device)

        def action(self, step, delta_pos, delta_rot):
            """the fingers close by 0.15 mm each over the first four frames (as Scene_interact.action does), then only move"""
            if step < 5:
                self.gripper.step(delta_pos, delta_rot, np.array([-1.5e-4]))
            else:
                self.gripper.step_simple(delta_pos, delta_rot)
            self.gripper.update_bound(self)
    return Scene(config3_state(N, **kw))


# ------------------------------------------------------------------------------------------------ strip partition
def strip_spec(R, M, rank, world, dx=0.002, dt=5e-3, seed=0, z0=0.0003, bump=0.25, noise=0.01, k_contact=40000.0, mu=0.5, ghost=2):
    """Rank `rank` of `world` of a sheet of world*R vertex rows x (M+1) columns cut into strips of R rows (R even keeps the
    alternating triangulation of Cloth.init_mesh aligned across strips): the local rows are the owned ones plus `ghost` rows on every
    inner side, expressed in a frame centred on the strip, over a frozen table slab that covers the strip.  world = 1 is the whole
    sheet on one GPU (the reference of the partition's parity test).  Start state as sheet_spec: flat + bump + seeded noise, defined on
    the GLOBAL grid so that every partition sees the same sheet."""
    assert R % 2 == 0 or world == 1, "strips must start on an even grid row"
    G = world * R
    g_lo = ghost if rank > 0 else 0
    g_hi = ghost if rank < world - 1 else 0
    a, b = rank * R - g_lo, (rank + 1) * R + g_hi                      # local rows [a, b) of the global grid
    rows = b - a
    rng = np.random.default_rng(seed)
    nz = rng.uniform(-noise * dx, noise * dx, (G, M + 1, 3))
    i, j = np.meshgrid(np.arange(a, b), np.arange(M + 1), indexing="ij")
    xg = i * dx - 0.5 * (G - 1) * dx
    xc = 0.5 * (xg[0, 0] + xg[-1, 0])                                    # local frame: strip centred at x = 0
    cpos = np.stack([xg - xc, j * dx - 0.5 * M * dx, z0 + bump * dx * (1 + np.sin(2 * np.pi * i / 32.0) * np.cos(2 * np.pi * j / 32.0))], -1)
    cpos = (cpos + nz[a:b]).reshape(-1, 3)
    sx, sy = (rows - 1) * dx + 0.02, M * dx + 0.02
    tdx = 0.003
    tnx, tny = int(np.ceil(sx / tdx)) + 1, int(np.ceil(sy / tdx)) + 1
    Len = tdx * (max(tnx, tny) - 1)
    toff = (-0.5 * (tnx - 1) * tdx, -0.5 * (tny - 1) * tdx, -tdx)
    grid_n = max(132, 2 * int(np.ceil((0.5 * max((tnx - 1) * tdx, (tny - 1) * tdx) + 0.01) / 0.003)) + 2)
    return dict(R=R, M=M, rank=rank, world=world, rows=rows, ghost_lo=g_lo, ghost_hi=g_hi, dx=dx, dt=dt, cloth_pos=cpos, x_shift=xc,
                table_size=Len, table_N=(tnx, tny, 2), table_offset=toff, grid_n=grid_n, k_contact=k_contact, mu=mu,
                max_n_constraints=rows * (M + 1) + 16, n_tris_owned=None, first_row_global=a, own0=g_lo * (M + 1), own1=(rows - g_hi) * (M + 1),
                n_tris_global=2 * (G - 1) * M)


def strip_scene(R, M, rank, world, device="cuda:0", **kw):
    """the local scene of one rank of the strip-partitioned sheet; engine.dist_init is called here (needs an initialised
    torch.distributed NCCL group when world > 1)"""
    import torch

    from .task_scene.Scene_bouncing import Scene
    sp = strip_spec(R, M, rank, world, **kw)
    rows = sp["rows"]
    s = Scene(cloth_size=(rows - 1) * sp["dx"], cloth_N=rows - 1, cloth_M=M, dt=sp["dt"], table_size=sp["table_size"], table_N=sp["table_N"],
              table_offset=sp["table_offset"], cloth_offset=(0.0, 0.0, 0.0), reset_offset=(0.0, 0.0, 0.0), k_contact=sp["k_contact"],
              max_n_constraints=sp["max_n_constraints"], grid_n=sp["grid_n"], device=device)
    s.mu_cloth_elastic[None] = sp["mu"]
    s.init_all()
    NV = rows * (M + 1)
    s.engine.pos[:NV] = torch.from_numpy(sp["cloth_pos"]).to(s.engine.device)
    s.engine.prev_pos.copy_(s.engine.pos)
    s.engine.vel.zero_()
    s.engine.cloth_ref_angle[0].zero_()
    s.engine.dist_init(rank, world, sp["ghost_lo"], sp["ghost_hi"], sp["first_row_global"])
    s.spec = sp
    return s
