"""analytic_grad_system.Grad on the B200 engine (code/engine/analytic_grad_system.py): trajectory buffers in torch CUDA
tensors, one tsl_step_backward call per transfer_grad."""
import torch

from ..fields import Scalar, TensorField


class Grad:
    def __init__(self, sys, tot_timestep, n_parts=0):
        e = sys.engine
        self.tot_NV, self.n_part, self.tot_timestep = sys.tot_NV, n_parts, tot_timestep
        f64 = dict(dtype=torch.float64, device=e.device)
        NF = sys.cloths[0].NF
        self._pos_buffer = torch.zeros((tot_timestep, sys.tot_NV, 3), **f64)
        self._pos_grad = torch.zeros((tot_timestep, sys.tot_NV, 3), **f64)
        self._ref_angle_buffer = torch.zeros((tot_timestep, sys.cloth_cnt, NF, 3), **f64)
        self._angleref_grad = torch.zeros((tot_timestep, sys.cloth_cnt, NF, 3), **f64)
        self._grad_kb = torch.zeros((1,), **f64)
        self._z = torch.zeros((3 * sys.tot_NV,), **f64)
        self.grad_kb = Scalar(0.0, lambda v: self._grad_kb.fill_(v), lambda: float(self._grad_kb.item()))
        self.grad_mu, self.grad_lam, self.grad_friction_coef = Scalar(), Scalar(), Scalar()
        self.dt, self.damping = sys.dt, 1.0
        self.count_friction_grad, self.count_mu_lam_grad, self.count_kb_grad = False, False, True
        self.clamp = 1.0                      # clamp_grad: +-1 (analytic_grad_system.py:104-108)
        self.last_solve = None

    pos_buffer = property(lambda self: TensorField(self._pos_buffer))
    pos_grad = property(lambda self: TensorField(self._pos_grad))
    ref_angle_buffer = property(lambda self: TensorField(self._ref_angle_buffer))
    angleref_grad = property(lambda self: TensorField(self._angleref_grad))

    def reset(self):
        self._pos_buffer.zero_(); self._pos_grad.zero_()
        self._grad_kb.zero_()
        # analytic_grad_system.py:33-39 also clears the other parameter gradients (and leaves angleref_grad untouched)
        self.grad_mu[None] = 0.0; self.grad_lam[None] = 0.0; self.grad_friction_coef[None] = 0.0

    def init_mass(self, sys):
        pass                                  # the engine reads sys.mass directly

    def copy_pos(self, sys, step):
        self._pos_buffer[step].copy_(sys.engine.pos)
        for k in range(self._ref_angle_buffer.shape[1]):
            self._ref_angle_buffer[step, k].copy_(sys.engine.cloth_ref_angle[k])

    def get_loss_table(self, sys):
        """analytic_grad_system.py:176-180 (row index uses cloth.N + 1, as the reference does)"""
        c = sys.cloths[0]
        row = (torch.arange(c.NV, device=self._pos_grad.device) / (c.N + 1)).to(torch.int64)
        sel = torch.nonzero((row == 5) | (row == 10)).flatten() + c.offset
        self._pos_grad[1:, sel, 2] = -1.0

    def get_loss_slide(self, sys, pos_grad=False):
        """:171-173"""
        c = sys.cloths[0]
        self._pos_grad[1:, c.offset:c.offset + c.NV, 0] = 1.0

    def get_loss_card(self, sys):
        """:175-177"""
        c = sys.cloths[0]
        self._pos_grad[self.tot_timestep - 1, c.offset:c.offset + c.NV, 0] = 1.0

    def transfer_grad(self, step, sys, f_contact=None, rel_tol=1e-10, max_iters=20000):
        pg_tm2 = self._pos_grad[step - 2] if step > 1 else None
        # (:147-152) with count_friction_grad the step feeds grad_friction_coef INSTEAD of the stiffness gradients
        kb_acc = self._grad_kb if (self.count_kb_grad and not self.count_friction_grad) else torch.zeros_like(self._grad_kb)
        self.last_solve = sys.engine.step_backward(
            self._pos_buffer[step], self._pos_buffer[step - 1], self._ref_angle_buffer[step - 1],        # (every cloth, one after the other)
            self._pos_grad[step], self._pos_grad[step - 1], pg_tm2, self._angleref_grad[step], self._angleref_grad[step - 1],
            kb_acc, self._z, clamp=self.clamp, rel_tol=rel_tol, max_iters=max_iters)
        if self.count_friction_grad:
            # Scene.contact_energy_backprop_friction (code/task_scene/Scene_sliding.py:140-177)
            # over the first nc1 constraints = the cloth-cloth pairs, which such a scene registers first; every pair otherwise
            self.grad_friction_coef[None] = self.grad_friction_coef[None] + sys.engine.friction_coef_grad(
                self._z, 0, getattr(sys, "n_cloth_cloth_pairs", None))
            return self.last_solve
        if self.count_mu_lam_grad and sys.engine.tet_bodies:
            # Grad.get_parameters_grad (:69-75): grad_mu / grad_lam += sum over free DOFs of z d_mu / z d_lam
            _, _, (gm, gl) = sys.engine.elastic_param_grad(self._z)
            self.grad_mu[None] = self.grad_mu[None] + gm
            self.grad_lam[None] = self.grad_lam[None] + gl
        return self.last_solve
