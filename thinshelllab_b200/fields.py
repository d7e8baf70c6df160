"""Tiny stand-ins for the Taichi field idioms the reference's driver scripts use on scene objects
(`x[None] = v`, `x[None]`, `.to_numpy()`, `.to_torch()`): thin views over Python floats / torch tensors."""
import torch


class Scalar:
    """`sys.cloths[0].Kb[None] = 120.0` / `analy_grad.grad_kb[None]`"""

    def __init__(self, value=0.0, on_set=None, getter=None):
        self._v, self._on_set, self._getter = float(value), on_set, getter

    def __getitem__(self, idx):
        return self._getter() if self._getter else self._v

    def __setitem__(self, idx, v):
        self._v = float(v)
        if self._on_set:
            self._on_set(self._v)


class TensorField:
    """read-mostly view of a torch tensor with the Taichi accessor names"""

    def __init__(self, t):
        self.t = t

    def to_numpy(self):
        return self.t.detach().cpu().numpy()

    def to_torch(self, device=None):
        return self.t.clone() if device is None else self.t.to(device)

    def from_numpy(self, a):
        self.t.copy_(torch.as_tensor(a, dtype=self.t.dtype).reshape(self.t.shape))

    def fill(self, v):
        self.t.fill_(v)

    def __getitem__(self, i):
        return self.t[i]

    def __setitem__(self, i, v):
        self.t[i] = torch.as_tensor(v, dtype=self.t.dtype, device=self.t.device)
