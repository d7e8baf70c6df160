"""agent_trajopt on torch tensors: the open-loop trajectory container of the reference's drivers
(code/agent/traj_opt_single.py:4-48).  O(T x parts) host-side bookkeeping, not a kernel target:
same attribute names and semantics so that driver scripts run unchanged."""
import torch

from ..fields import TensorField


class agent_trajopt:
    def __init__(self, tot_timestep, cnt, max_moving_dist=0.0005):
        self._traj = torch.zeros((tot_timestep, cnt, 6), dtype=torch.float64)
        self._delta_pos = torch.zeros((cnt, 3), dtype=torch.float64)
        self._delta_rot = torch.zeros((cnt, 3), dtype=torch.float64)
        self.tmp_action = TensorField(torch.zeros((cnt, 6), dtype=torch.float64))
        self.action_dim = 6 * cnt
        self.tot_timestep, self.max_moving_dist, self.n_part = tot_timestep, max_moving_dist, cnt

    traj = property(lambda self: TensorField(self._traj))
    delta_pos = property(lambda self: TensorField(self._delta_pos))
    delta_rot = property(lambda self: TensorField(self._delta_rot))

    def calculate_dist(self, frame, max_dist, j):
        d = self._traj[frame, j] - self._traj[frame - 1, j]
        return float(d[:3].norm() + d[3:].norm() * max_dist)

    def fix_action(self, max_dist):
        """rescales every per-frame increment so that |dpos| + max_dist |drot| <= max_moving_dist (sequentially in time, :15-27)"""
        for i in range(1, self.tot_timestep):
            for j in range(self.n_part):
                d = self._traj[i, j] - self._traj[i - 1, j]
                w = self.max_moving_dist / (float(d[:3].norm() + d[3:].norm() * max_dist) + 1e-8)
                if w < 1.0:
                    self._traj[i, j] = self._traj[i - 1, j] + d * w

    def init_traj_pick_fold(self):
        """:62-73: both pads press down 0.6 mm per frame for 8 frames, then hold"""
        for i in range(min(8, self.tot_timestep)):
            self._traj[i, 0, 2] = -0.0006 * i
            self._traj[i, 1, 2] = -0.0006 * i
            self._traj[i, 0, 0] = self._traj[i - 1, 0, 0]
            self._traj[i, 1, 0] = self._traj[i - 1, 1, 0]
        for i in range(8, min(50, self.tot_timestep)):
            self._traj[i, :2, 2] = self._traj[i - 1, :2, 2]
            self._traj[i, :2, 0] = self._traj[i - 1, :2, 0]

    def init_traj_card(self):
        """:75-96: the two end pads close on the stack, then pad 0 lifts and (from frame 35) turns its end"""
        t, T = self._traj, self.tot_timestep
        for i in range(min(5, T)):
            t[i, 0, 0] = t[i - 1, 0, 0] + 0.0003
            t[i, 1, 0] = t[i - 1, 1, 0] - 0.0003
        for lo, hi, dx, dz, dr in ((5, 20, 0.0001, 0.0003, 0.0), (20, 35, 0.0001, 0.0002, 0.0), (35, 50, 0.0002, 0.0005, 0.02)):
            for i in range(lo, min(hi, T)):
                t[i, 0, 0] = t[i - 1, 0, 0] + dx
                t[i, 0, 2] = t[i - 1, 0, 2] + dz
                t[i, 0, 4] = t[i - 1, 0, 4] + dr
                t[i, 1, 0] = t[i - 1, 1, 0]

    def init_traj_slide(self):
        """:104-109: press down for 10 frames, then drag along -x"""
        t, T = self._traj, self.tot_timestep
        for i in range(min(10, T)):
            t[i, 0, 2] = -0.00035 * i
        for i in range(10, min(50, T)):
            t[i, 0, 0] = t[i - 1, 0, 0] - 0.0005
            t[i, 0, 2] = t[i - 1, 0, 2]

    def get_action(self, step):
        d = self._traj[step] - self._traj[step - 1]
        self._delta_pos.copy_(d[:, :3])
        self._delta_rot.copy_(d[:, 3:])
