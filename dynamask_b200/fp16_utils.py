"""Mixed-precision guard used by the RoI extractors.

Same contract as the reference decorator (``mmdet/core/fp16/decorators.py:88-164``): when the
owning module has ``fp16_enabled`` set, the named tensor arguments are cast half -> float before
the call and (``out_fp16=True``) the result is cast float -> half afterwards; otherwise the call
is untouched.  The kernels therefore only ever see fp32.
"""
import functools
from inspect import getfullargspec

import torch


def _cast(x, src, dst):
    if isinstance(x, torch.Tensor):
        return x.to(dst) if x.dtype == src else x
    if isinstance(x, (list, tuple)):
        return type(x)(_cast(v, src, dst) for v in x)
    if isinstance(x, dict):
        return {k: _cast(v, src, dst) for k, v in x.items()}
    return x


def force_fp32(apply_to=None, out_fp16=False):
    def wrap(fn):
        spec = getfullargspec(fn)

        @functools.wraps(fn)
        def inner(*args, **kwargs):
            if not isinstance(args[0], torch.nn.Module):
                raise TypeError('@force_fp32 can only be used to decorate the method of nn.Module')
            if not getattr(args[0], 'fp16_enabled', False):
                return fn(*args, **kwargs)
            names = spec.args if apply_to is None else apply_to
            new_args = [
                _cast(a, torch.half, torch.float) if n in names else a
                for n, a in zip(spec.args[:len(args)], args)
            ]
            new_kwargs = {
                k: (_cast(v, torch.half, torch.float) if k in names else v)
                for k, v in kwargs.items()
            }
            out = fn(*new_args, **new_kwargs)
            return _cast(out, torch.float, torch.half) if out_fp16 else out

        return inner

    return wrap
