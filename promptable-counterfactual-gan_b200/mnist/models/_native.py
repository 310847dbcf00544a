"""Shared plumbing of the three mirror modules: arena binding + forward-only native plans."""
import torch
import torch.nn as nn

from .. import binding
from .. import plan as P


class NativeNet(nn.Module):
    """Base of the mirror modules.  Sub-classes set ``_net`` (0 generator, 1 discriminator, 2 classifier)."""
    _net = -1
    precision = "bf16"        # "fp32" selects the exact CUDA-core mode

    def _plan_for(self, batch):
        dev = next(self.parameters()).device
        arena = binding.bind(self, self._net, dev)
        cache = self.__dict__.setdefault("_pcg_plans", {})
        key = (batch, self.precision, tuple(getattr(self, "_pcg_ptrs", ())))
        plan = cache.get(key)
        if plan is None:
            for old in cache.values():
                old.close()
            cache.clear()
            if self._net == 0:
                base_ch, n_res = self._pcg_dims
                running, nbt = self._pcg_bn
                g = arena
            else:
                base_ch, n_res = 4, 1
                g = binding.dummy_arena(0, dev, base_ch, n_res)
                running = torch.zeros(2 * n_res, 2, base_ch, device=dev)
                nbt = torch.zeros(2 * n_res, dtype=torch.int64, device=dev)
            d = arena if self._net == 1 else binding.dummy_arena(1, dev)
            c = arena if self._net == 2 else binding.dummy_arena(2, dev)
            plan = P.MnistStepPlan(batch, g, d, c, running, nbt, P.StepConfig(precision=self.precision),
                                   base_ch, n_res)
            cache[key] = plan
            self.__dict__["_pcg_seen"] = None
        ver = binding.param_versions(self)
        if self.__dict__.get("_pcg_seen") != ver:      # parameters were written since the last packing
            plan.refresh_weights()
            self.__dict__["_pcg_seen"] = ver
        return plan

    @staticmethod
    def _img(x):
        if not x.is_cuda:
            raise RuntimeError("pcg_b200: inputs must be CUDA tensors (there is no CPU fallback)")
        return x.detach().to(torch.float32).contiguous()
