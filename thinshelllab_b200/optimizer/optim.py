"""Adam_single on torch tensors (code/optimizer/optim.py:37-81): epsilon inside the square root, learning rate multiplied by
`discount` every 10 steps.  Host-side, O(T x parts x 6)."""
import torch


def _t(x):
    return x.t if hasattr(x, "t") and isinstance(getattr(x, "t"), torch.Tensor) else x


class Adam_single:
    def __init__(self, parameters_shape, lr, beta_1, beta_2, eps, discount=0.9):
        self.shape = tuple(parameters_shape)
        self.beta_1, self.beta_2, self.eps, self.discount = float(beta_1), float(beta_2), float(eps), discount
        self.momentum_buffer = torch.zeros(self.shape, dtype=torch.float64)
        self.v_buffer = torch.zeros(self.shape, dtype=torch.float64)
        self.iter, self.lr, self.ori_lr = 0.0, float(lr), float(lr)

    def step(self, parameters, grads):
        p, g = _t(parameters), _t(grads).to(torch.float64).cpu()
        if torch.isnan(g).any():
            print("nan in gripper grid!!")
        self.momentum_buffer = self.beta_1 * self.momentum_buffer + (1 - self.beta_1) * g
        self.v_buffer = self.beta_2 * self.v_buffer + (1 - self.beta_2) * g * g
        m_cap = self.momentum_buffer / (1 - self.beta_1 ** (self.iter + 1))
        v_cap = self.v_buffer / (1 - self.beta_2 ** (self.iter + 1))
        p -= ((self.lr * m_cap) / torch.sqrt(v_cap + self.eps)).to(p.device)
        self.iter += 1.0
        if int(self.iter) % 10 == 0:
            self.lr *= self.discount

    def reset(self):
        self.iter, self.lr = 0.0, self.ori_lr
        self.momentum_buffer.zero_(); self.v_buffer.zero_()
