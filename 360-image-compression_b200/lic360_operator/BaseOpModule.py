"""BaseOpModule -- common base of the operator modules (mirrors lic360_operator/BaseOpModule.py:5-55).

Each module owns `self.op = {gpu_id: lic360.<X>Op(...)}` and dispatches on `x.device.index`.  Moving the module
(`.to(dev)`, `.cuda(i)`) re-keys that dict and re-binds the native op, which is what the reference achieves by
re-entering `to()` from `_apply` (BaseOpModule.py:12-16,33-55).  `nn.DataParallel` replicas share the dict because
`nn.Module._replicate_for_data_parallel` copies `__dict__` shallowly (BaseOpModule.py:22-31).
"""
import torch
from torch import nn


class BaseOpModule(nn.Module):

    def __init__(self, devices=0):
        super(BaseOpModule, self).__init__()
        self.device_list = [devices] if isinstance(devices, int) else list(devices)
        self.apply_flag = False

    def _apply(self, fn, *args, **kwargs):
        super(BaseOpModule, self)._apply(fn, *args, **kwargs)
        try:
            target = fn(torch.empty(0)).device
        except Exception:  # fn not applicable to a probe tensor: nothing to re-bind
            return self
        if target.type == 'cuda':
            self.custom_op_to(target)
        return self

    def custom_op_to(self, *args):
        dev = args[0] if args else None
        ops = getattr(self, 'op', None)
        if dev is None or not ops or len(ops) != 1:
            return
        new_id = dev.index if dev.index is not None else torch.cuda.current_device()
        old_id = next(iter(ops))
        if new_id != old_id:
            ops[new_id] = ops.pop(old_id)
            ops[new_id].to(new_id)

    def custom_op_replicate(self, other):
        other.op = self.op
        return other
