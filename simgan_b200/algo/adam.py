"""Adam state for the fused update kernels, with the slice of the torch.optim surface the reference
touches: ``param_groups[i]['lr']`` is read at every update (third_party/a2c_ppo_acktr/utils.py:68-72
mutates it), ``zero_grad()``/``state_dict()`` exist.  The arithmetic itself runs inside the CUDA
kernels (sg_common.cuh: adam_update) in the op order of torch.optim.Adam's single-tensor path."""
import numpy as np
import torch


class FusedAdam(object):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.param_groups = [dict(params=list(params), lr=lr, betas=tuple(betas), eps=eps, weight_decay=0,
                                  amsgrad=False)]
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps)
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None
        self._pending = None        # per-parameter moments imported from a torch.optim.Adam (from_torch_adam)

    @classmethod
    def from_torch_adam(cls, opt, params):
        """FusedAdam equivalent of a ``torch.optim.Adam`` over ``params`` (a reference checkpoint's optimizer):
        lr / betas / eps of its first group and, when it has stepped, exp_avg / exp_avg_sq / step of every parameter.
        The moments are folded into the flat buffers at the first ``ensure_state``."""
        g = opt.param_groups[0]
        new = cls(params, lr=g["lr"], betas=g["betas"], eps=g["eps"])
        moments, step = [], 0
        for p_old in g["params"]:
            st = opt.state.get(p_old, {})
            if "exp_avg" not in st:
                moments = None
                break
            moments.append((st["exp_avg"].detach().clone(), st["exp_avg_sq"].detach().clone()))
            step = int(st["step"]) if not torch.is_tensor(st["step"]) else int(st["step"].item())
        if moments:
            new._pending = (moments, step)
        return new

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    def zero_grad(self, set_to_none=True):
        return None

    def ensure_state(self, flat):
        """Allocate (or re-allocate after a device / size change) the flat moment buffers."""
        if self.exp_avg is None or self.exp_avg.numel() != flat.numel() or self.exp_avg.device != flat.device:
            self.exp_avg = torch.zeros_like(flat)
            self.exp_avg_sq = torch.zeros_like(flat)
            self.step_count = 0
            pending, self._pending = getattr(self, "_pending", None), None
            if pending is not None:
                moments, step = pending
                for p, (m, v) in zip(self.param_groups[0]["params"], moments):
                    off = (p.data_ptr() - flat.data_ptr()) // 4        # the parameters are views of the flat vector
                    assert 0 <= off and off + p.numel() <= flat.numel(), "parameters are not bound to the flat vector"
                    self.exp_avg[off:off + p.numel()].copy_(m.reshape(-1))
                    self.exp_avg_sq[off:off + p.numel()].copy_(v.reshape(-1))
                self.step_count = step
        return self.exp_avg, self.exp_avg_sq

    def schedule(self, n_steps):
        """Per-step scalars torch forms in Python double: lr/(1-beta1^t) and sqrt(1-beta2^t),
        t = step_count+1 .. step_count+n_steps, narrowed to fp32 like a Scalar operand."""
        g = self.param_groups[0]
        lr, (b1, b2) = g["lr"], g["betas"]
        out = np.empty((2, n_steps), dtype=np.float32)
        for i in range(n_steps):
            t = self.step_count + 1 + i
            out[0, i] = lr / (1 - b1 ** t)
            out[1, i] = (1 - b2 ** t) ** 0.5
        return out

    def state_dict(self):
        return dict(step=self.step_count, exp_avg=self.exp_avg, exp_avg_sq=self.exp_avg_sq,
                    param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups])

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg, self.exp_avg_sq = sd["exp_avg"], sd["exp_avg_sq"]
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
