"""nerf.TinyNeRF -- drop-in for the reference's src/nerf.py:4-41."""
import torch
import torch.nn as nn

import _engine as E
import _lazy


class _MLP(torch.autograd.Function):
    """fp32 stand-alone MLP (tnerf_mlp_fwd / tnerf_mlp_bwd).  Inputs: x, then the parameters in state_dict order."""

    @staticmethod
    def forward(ctx, module, x, *params):
        dev = E.need_cuda(x)
        h = E.handle_for(module, dev)
        h.bind()
        xc = E.f32c(x)
        n = xc.shape[0]
        rgb = torch.empty((n, 3), dtype=torch.float32, device=dev)
        sigma = torch.empty((n, 1), dtype=torch.float32, device=dev)
        need_grad = any(ctx.needs_input_grad)     # (grad mode is always off inside Function.forward)
        acts = torch.empty((module.depth, n, module.hidden), dtype=torch.float32, device=dev) if need_grad else None
        E.check(E.lib().tnerf_mlp_fwd(h.h, E.ptr(xc), n, E.ptr(rgb), E.ptr(sigma), E.ptr(acts), E.stream(dev)), "tnerf_mlp_fwd")
        if need_grad:
            ctx.save_for_backward(xc, acts, rgb, sigma, *params)      # parameters: guarded by autograd's in-place-modification check
            ctx.module = module
            ctx.x_grad = x.requires_grad
        return rgb, sigma

    @staticmethod
    def backward(ctx, g_rgb, g_sigma):
        xc, acts, rgb, sigma = ctx.saved_tensors[:4]
        module, dev, n = ctx.module, xc.device, xc.shape[0]
        h = E.handle_for(module, dev)
        ps = h.bind()
        grads = torch.zeros(h.param_count, dtype=torch.float32, device=dev)
        gx = torch.empty_like(xc) if ctx.x_grad else None
        scratch = torch.empty(int(E.lib().tnerf_mlp_bwd_scratch_floats(h.h, n)), dtype=torch.float32, device=dev)
        g_rgb, g_sigma = E.f32c(g_rgb), E.f32c(g_sigma)     # keep alive until the launch is enqueued
        E.check(E.lib().tnerf_mlp_bwd(h.h, E.ptr(xc), n, E.ptr(acts), E.ptr(rgb), E.ptr(sigma), E.ptr(g_rgb),
                                      E.ptr(g_sigma), E.ptr(grads), E.ptr(gx), E.ptr(scratch), E.stream(dev)),
                "tnerf_mlp_bwd")
        views = E.flat_grad_views(module, grads)
        return (None, gx) + tuple(v if p.requires_grad else None for v, p in zip(views, ps))


class TinyNeRF(nn.Module):
    """NeRF-style MLP without view directions: ``depth`` Linear(hidden)+ReLU layers, the encoded input
    concatenated BEHIND the activations after layer ``skip_at-1``, heads rgb = sigmoid(Linear(hidden,3))
    and sigma = relu(Linear(hidden,1)).  Parameters are ordinary fp32 ``nn.Linear`` parameters (same
    names, shapes and construction order as the reference, so seeds and checkpoints interchange)."""

    def __init__(self, in_dim: int, hidden: int = 128, depth: int = 4, skip_at: int = 2):
        super().__init__()
        self.in_dim, self.hidden, self.depth, self.skip_at = in_dim, hidden, depth, skip_at
        self.layers = nn.ModuleList()
        width = in_dim
        for i in range(depth):
            self.layers.append(nn.Linear(width, hidden))
            width = hidden + in_dim if i == skip_at - 1 else hidden
        self.sigma = nn.Sequential(nn.Linear(hidden, 1), nn.ReLU(inplace=True))
        self.rgb = nn.Sequential(nn.Linear(hidden, 3), nn.Sigmoid())

    def _params(self):
        ps = []
        for lin in self.layers:
            ps += [lin.weight, lin.bias]
        return ps + [self.sigma[0].weight, self.sigma[0].bias, self.rgb[0].weight, self.rgb[0].bias]

    def _forward_dense(self, x: torch.Tensor):
        if self.skip_at == self.depth:
            raise RuntimeError(f"mat1 and mat2 shapes cannot be multiplied: heads expect {self.hidden} features but "
                               f"skip_at == depth feeds {self.hidden + self.in_dim}")
        if x.dim() != 2 or x.shape[1] != self.in_dim:
            raise RuntimeError(f"TinyNeRF expects (N, {self.in_dim}), got {tuple(x.shape)}")
        return _MLP.apply(self, x, *self._params())

    def forward(self, x):
        """x (N, in_dim) encoded coordinates -> rgb (N,3) in (0,1), sigma (N,1) >= 0."""
        if (isinstance(x, _lazy.Deferred) and x._kind == "enc" and x._node.model is None and x.dim() == 2
                and x.shape[1] == self.in_dim and x._node.encoder.out_dim == self.in_dim and self.skip_at != self.depth):
            x._node.model = self
            n = x.shape[0]
            return (_lazy.Deferred((n, 3), x.device, x._node, "rgb"), _lazy.Deferred((n, 1), x.device, x._node, "sigma"))
        return self._forward_dense(_lazy._real(x))
