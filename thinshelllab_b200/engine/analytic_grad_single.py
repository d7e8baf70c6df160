"""analytic_grad_single.Grad on the B200 engine (code/engine/analytic_grad_single.py): the adjoint of a rollout with respect to
the gripper trajectory.  Trajectory buffers are torch CUDA tensors; one tsl_step_backward_ex call per transfer_grad (adjoint solve,
counting pass for tmp_z_frozen, friction / rest-angle lag terms, time recurrence) plus the gather over the driven vertices."""
import numpy as np
import torch

from ..fields import TensorField


class Grad:
    def __init__(self, sys, tot_timestep, n_parts, friction_loss=False, f_loss_ratio=0.001, vertical_only=False):
        e = sys.engine
        self.tot_NV, self.n_part, self.tot_timestep = sys.tot_NV, n_parts, tot_timestep
        f64 = dict(dtype=torch.float64, device=e.device)
        NF = sys.cloths[0].NF
        self.NF, self.cloth_cnt = NF, sys.cloth_cnt
        self._pos_buffer = torch.zeros((tot_timestep, sys.tot_NV, 3), **f64)
        self._pos_grad = torch.zeros((tot_timestep, sys.tot_NV, 3), **f64)
        self._ref_angle_buffer = torch.zeros((tot_timestep, sys.cloth_cnt, NF, 3), **f64)
        self._angleref_grad = torch.zeros((tot_timestep, sys.cloth_cnt, NF, 3), **f64)
        self._gripper_pos_buffer = np.zeros((tot_timestep, n_parts, 3))
        self._gripper_rot_buffer = np.zeros((tot_timestep, n_parts, 4))
        self._gripper_grad = np.zeros((tot_timestep, n_parts, 6))
        self._z = torch.zeros((3 * sys.tot_NV,), **f64)
        self._z_frozen = torch.zeros((3 * sys.tot_NV,), **f64)
        self.dt, self.damping = sys.dt, 1.0
        self.friction_loss, self.f_loss_ratio, self.vertical_only = friction_loss, f_loss_ratio, vertical_only
        self.clamp = 1000.0                   # clamp_grad (analytic_grad_single.py)
        self.last_solve = None

    pos_buffer = property(lambda self: TensorField(self._pos_buffer))
    pos_grad = property(lambda self: TensorField(self._pos_grad))
    ref_angle_buffer = property(lambda self: TensorField(self._ref_angle_buffer))
    angleref_grad = property(lambda self: TensorField(self._angleref_grad))
    gripper_pos_buffer = property(lambda self: TensorField(torch.from_numpy(self._gripper_pos_buffer)))
    gripper_rot_buffer = property(lambda self: TensorField(torch.from_numpy(self._gripper_rot_buffer)))
    gripper_grad = property(lambda self: TensorField(torch.from_numpy(self._gripper_grad)))
    tmp_z_frozen = property(lambda self: TensorField(self._z_frozen))

    def reset(self):
        self._pos_buffer.zero_(); self._pos_grad.zero_(); self._angleref_grad.zero_()

    def init_mass(self, sys):
        pass                                  # the engine reads sys.mass directly

    def copy_pos(self, sys, step):
        """:38-51"""
        self._pos_buffer[step].copy_(sys.engine.pos)
        for k in range(self._ref_angle_buffer.shape[1]):
            self._ref_angle_buffer[step, k].copy_(sys.engine.cloth_ref_angle[k])
        self._gripper_pos_buffer[step] = sys.gripper._pos
        self._gripper_rot_buffer[step] = sys.gripper._rot

    def get_loss_fold(self, sys, curve7, curve8):
        """seed on the rest angles of the two crease rows of the last frame (analytic_grad_single.py:281-292; the loss of
        training/trajopt_folding.py:130)"""
        sel7, sel8 = sys._crease_hinges()
        g = self._angleref_grad[self.tot_timestep - 1, 0]
        g[sel7[0], sel7[1]] = curve7
        g[sel8[0], sel8[1]] = curve8

    def get_loss_push(self, sys, target_pos):
        """:297-300: d/dx of the squared distance of the last frame's cloth to the target shape"""
        c = sys.cloths[0]
        t = torch.as_tensor(np.asarray(target_pos, np.float64), device=self._pos_grad.device)
        T = self.tot_timestep
        self._pos_grad[T - 1, c.offset:c.offset + c.NV] = 2.0 * (self._pos_buffer[T - 1, c.offset:c.offset + c.NV] - t)

    # ---- the remaining loss seeds of analytic_grad_single.py:258-471 that apply to one-cloth scenes (plain writes into pos_grad /
    # angleref_grad, as the reference's kernels)
    def _cloth(self, sys):
        c = sys.cloths[0]
        return c, slice(c.offset, c.offset + c.NV)

    def get_loss(self, sys):
        """:258-262"""
        c, sl = self._cloth(sys)
        self._pos_grad[:, :c.NV, 0] = -1.0              # (the reference indexes the first NV rows, not offset + i)

    def get_loss_sheet(self, sys):
        """:264-268"""
        c, sl = self._cloth(sys)
        self._pos_grad[1:, :c.NV, 0] = 1.0

    def get_loss_book(self, sys):
        """:273-277"""
        c, sl = self._cloth(sys)
        self._pos_grad[1:, :c.NV, 0] = -1.0

    def _row_is(self, sys, row):
        c = sys.cloths[0]
        return torch.nonzero((torch.arange(c.NV, device=self._pos_grad.device) // (c.M + 1)) == row).flatten() + c.offset

    def get_loss_pick(self, sys):
        """:324-327 (get_loss_card :384-388 is the same seed): lift grid row 8"""
        self._pos_grad[:, self._row_is(sys, 8), 2] = -1.0

    get_loss_card = get_loss_pick

    def get_loss_pick_fold(self, sys):
        """:373-382: rest angles of the crease between grid rows 7 and 9, every frame"""
        _, s8 = sys._crease_hinges()
        self._angleref_grad[:, 0, s8[0], s8[1]] = -1.0

    def get_loss_slide_simple(self, sys):
        """:390-393"""
        c, sl = self._cloth(sys)
        self._pos_grad[self.tot_timestep - 1, sl, 0] = 1.0

    def get_loss_interact(self, sys):
        """:408-420"""
        c, sl = self._cloth(sys)
        b = sys.elastics[3]
        self._pos_grad[self.tot_timestep - 1, sl, 0] = 1.0
        self._pos_grad[self.tot_timestep - 1, b.offset:b.offset + b.n_verts, 0] = -256.0 / 144.0

    def get_loss_interact_1(self, sys):
        """:422-426"""
        b = sys.elastics[3]
        self._pos_grad[self.tot_timestep - 1, b.offset:b.offset + b.n_verts, 0] = 1.0

    def _loss_towards_cloth_vertex(self, sys, tt):
        """get_loss_balance / get_loss_side (:428-463): the carried body elastics[0] follows cloth vertex tt in x, y on every frame
        (the cloth-vertex entries are plain assignments in the reference's parallel loop: the last body vertex wins)"""
        c = sys.cloths[0]
        b = sys.elastics[0]
        pb = self._pos_buffer
        for j in range(1, self.tot_timestep):
            d = pb[j, b.offset:b.offset + b.n_verts, :2] - pb[j, c.offset + tt, :2]
            self._pos_grad[j, b.offset:b.offset + b.n_verts, :2] = 2 * d
            self._pos_grad[j, c.offset + tt, :2] = -2 * d[-1]

    def get_loss_balance(self, sys):
        self._loss_towards_cloth_vertex(sys, (sys.cloth_N + 1) // 2 * (sys.cloth_M + 1) + (sys.cloth_M + 1) // 2)

    def get_loss_side(self, sys):
        self._loss_towards_cloth_vertex(sys, (sys.cloth_N + 1) // 4 * (sys.cloth_M + 1) + (sys.cloth_M + 1) // 2)

    def get_loss_throwing(self, sys):
        """:465-473"""
        c = sys.cloths[0]
        b = sys.elastics[0]
        M = sys.cloth_M
        self._pos_grad[1:, b.offset:b.offset + b.n_verts, 2] = -1.0
        first = torch.arange(M, device=self._pos_grad.device) + c.offset
        last = first + sys.cloth_N * (M + 1)
        for rows in (first, last):
            self._pos_grad[1:, rows, 2] = 20 * self._pos_buffer[1:, rows, 2]

    def accumulate_gripper_grad(self, traj, max_dist):
        """:491-502"""
        for step in range(self.tot_timestep - 2, 1, -1):
            for j in range(self.n_part):
                if traj.calculate_dist(step + 1, max_dist, j) > traj.max_moving_dist - 0.00005:
                    self._gripper_grad[step, j] += self._gripper_grad[step + 1, j]

    def get_loss_lift(self, sys):
        """:303-312: d/dx of the squared distance of the carried box (elastics[0]) from its first-frame shape shifted by (-0.012, -0.012, 0)"""
        b = sys.elastics[0]
        T = self.tot_timestep
        d = self._pos_buffer[T - 1, b.offset:b.offset + b.n_verts] - self._pos_buffer[0, b.offset:b.offset + b.n_verts]
        d[:, 0] += 0.012; d[:, 1] += 0.012
        self._pos_grad[T - 1, b.offset:b.offset + b.n_verts] = d

    def apply_action_limit_grad(self, traj, max_dist):
        """:504-516: penalty gradient on steps whose pose increment exceeds the agent's max_moving_dist"""
        tr = traj.traj.to_numpy()
        for step in range(1, self.tot_timestep):
            for j in range(self.n_part):
                dist = traj.calculate_dist(step, max_dist, j)
                if dist > traj.max_moving_dist:
                    d = tr[step, j] - tr[step - 1, j]
                    self._gripper_grad[step, j, :3] += d[:3] * (dist - traj.max_moving_dist) * 10000000
                    self._gripper_grad[step, j, 3:] += d[3:] * (dist - traj.max_moving_dist) * 100000

    def transfer_grad(self, step, sys, f_contact=None, rel_tol=1e-10, max_iters=20000):
        pg_tm2 = self._pos_grad[step - 2] if step > 1 else None
        self.last_solve = sys.engine.step_backward_ex(
            self._pos_buffer[step], self._pos_buffer[step - 1], self._ref_angle_buffer[step - 1],         # (every cloth, one after the other)
            self._pos_grad[step], self._pos_grad[step - 1], pg_tm2, self._angleref_grad[step], self._angleref_grad[step - 1],
            None, self._z, self._z_frozen, clamp=self.clamp, clamp_angleref=self.clamp, rel_tol=rel_tol, max_iters=max_iters)
        if step > 0:
            self.get_gripper_grad(step, sys)
        return self.last_solve

    def get_gripper_grad(self, step, sys):
        """:118-135: pose of frame `step`, gather over the driven vertices"""
        sys.gripper.set(self._gripper_pos_buffer, self._gripper_rot_buffer, step)
        sys.gripper.get_rotmat()
        sys.gripper.gather_grad(self._z_frozen, sys)
        for j in range(self.n_part):
            if self.vertical_only:
                self._gripper_grad[step, j, 2] = sys.gripper._d_pos[j][2]
            else:
                self._gripper_grad[step, j, :3] = sys.gripper._d_pos[j]
                self._gripper_grad[step, j, 3:] = sys.gripper._d_angle[j]
