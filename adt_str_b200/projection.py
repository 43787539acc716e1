"""``ProjectToMel``: the reference's ``ADTModel.project_to_mel`` (``nn.Linear(n_mels, d_query * nhead)``,
``model.py:224-226``, applied to the log-mel at ``:249`` under bf16 autocast) on the tcgen05 tensor cores
(``csrc/project.cu``).

Same parameters as the ``nn.Linear`` it replaces (``weight`` (n_out, n_mels), ``bias`` (n_out,)), so a checkpoint's
``project_to_mel.weight`` / ``.bias`` load unchanged.  ``forward`` takes the float32 log-mel ``(..., n_mels)`` and returns
bfloat16 ``(..., n_out)`` - what ``F.linear`` gives under ``torch.autocast("cuda", torch.bfloat16)``: inputs, weight and
bias rounded to bf16, float32 accumulation, bf16 result.  The weight image on the device is rebuilt when the
parameters change (training).  Gradients: ``weight`` and ``bias`` get theirs through plain library GEMMs (the input is
data - the log-mel has no trainable ancestor - so no input gradient is produced unless asked for).
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib


class _Native:
    def __init__(self, weight: torch.Tensor, bias, device: torch.device):
        self.lib = _lib.load()
        _lib.check(self.lib.adtfe_device_ok(device.index or 0), "adtfe_device_ok")
        w = weight.detach().to("cpu", torch.float32).contiguous()
        b = None if bias is None else bias.detach().to("cpu", torch.float32).contiguous()
        h = C.c_void_p()
        _lib.check(self.lib.adtfe_linear_create(w.shape[1], w.shape[0], w.data_ptr(), None if b is None else b.data_ptr(),
                                                device.index or 0, C.byref(h)), "adtfe_linear_create")
        self.handle = h

    def __del__(self):
        try:
            if self.handle:
                self.lib.adtfe_linear_destroy(self.handle)
        except Exception:
            pass


class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, module):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return module._run(x)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        g = grad_out.reshape(-1, grad_out.shape[-1])
        x2 = x.reshape(-1, x.shape[-1])
        grad_x = (g.to(weight.dtype) @ weight).reshape(x.shape).to(x.dtype) if ctx.needs_input_grad[0] else None
        grad_w = (g.float().t() @ x2.float()).to(weight.dtype) if ctx.needs_input_grad[1] else None
        grad_b = g.float().sum(0) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return grad_x, grad_w, grad_b, None


class ProjectToMel(nn.Module):
    def __init__(self, in_features: int = 128, out_features: int = 768, bias: bool = True):
        super().__init__()
        ref = nn.Linear(in_features, out_features, bias=bias)      # the reference's initialisation (model.py:224-226)
        self.in_features, self.out_features = in_features, out_features
        self.weight = ref.weight
        self.bias = ref.bias
        self._native = None
        self._version = None

    @classmethod
    def from_linear(cls, linear: nn.Linear) -> "ProjectToMel":
        m = cls(linear.in_features, linear.out_features, linear.bias is not None)
        m.weight, m.bias = linear.weight, linear.bias
        return m

    def _handle(self, device: torch.device) -> _Native:
        version = (self.weight._version, None if self.bias is None else self.bias._version, str(device),
                   self.weight.data_ptr())
        if self._native is None or self._version != version:
            self._native = _Native(self.weight, self.bias, device)
            self._version = version
        return self._native

    def _run(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError("ProjectToMel needs a CUDA device (sm_100a); there is no CPU path")
        rows = x.reshape(-1, self.in_features).float().contiguous()
        out = torch.empty((rows.shape[0], self.out_features), dtype=torch.bfloat16, device=x.device)
        with torch.cuda.device(x.device):
            native = self._handle(x.device)
            _lib.check(native.lib.adtfe_linear_forward(native.handle, rows.data_ptr(), rows.shape[0], out.data_ptr(),
                                                       torch.cuda.current_stream(x.device).cuda_stream),
                       "adtfe_linear_forward")
        return out.reshape(*x.shape[:-1], self.out_features)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_grad_enabled() and (self.weight.requires_grad or x.requires_grad):
            return _Project.apply(x, self.weight, self.bias, self)
        return self._run(x)
