"""V2ce3d (stage 1) -- torch fp32 CPU restatement of the 3D-UNet forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the floating-point
reference the CUDA conv path is compared with (tolerances in the tests).  It
restates, in "folded" form (conv -> per-channel scale/shift -> activation), the
same network the reference builds:

  forward            <- /root/reference/scripts/v2ce_3d.py:26-30, scripts/unet_2layer.py:335-379
  residual block     <- /root/reference/scripts/submodules.py:216-264 (the 1x1x1 conv+BN shortcut
                        exists on EVERY block)
  head / pred        <- /root/reference/scripts/submodules.py:85-124, unet_2layer.py:235-236,291-297
  spectral norm step <- /root/reference/scripts/spectral_norm.py:19-31,62-64 (runs on every
                        forward, also in eval mode: SURVEY.md F3)
  nearest upsample   <- /root/reference/scripts/unet_2layer.py:358-364 (src = floor(dst*in/out))

State-dict key layout is the reference's (``UNet.encoders.0.conv1.weight``,
``UNet.decoders.0.conv1.module.weight_bar`` ...).
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def sn_conv_names():
    names = []
    for i in range(2):
        names += [f'UNet.resblocks.{i}.conv1', f'UNet.resblocks.{i}.conv2']
    for i in range(4):
        names += [f'UNet.decoders.{i}.conv1', f'UNet.decoders.{i}.conv2']
    return names


def l2n(v, eps=1e-12):
    return v / (v.norm() + eps)


class UNetOracle:
    def __init__(self, state_dict):
        self.sd = {k: v.detach().clone().float() for k, v in state_dict.items()}
        self.calls = 0

    # -- spectral norm -----------------------------------------------------
    def sn_step(self):
        """One power iteration per SN conv; returns the 12 sigmas of this call."""
        sig = []
        for name in sn_conv_names():
            w = self.sd[name + '.module.weight_bar']
            u = self.sd[name + '.module.weight_u']
            v = self.sd[name + '.module.weight_v']
            wm = w.reshape(w.shape[0], -1)
            v = l2n(torch.mv(wm.t(), u))
            u = l2n(torch.mv(wm, v))
            sigma = torch.dot(u, torch.mv(wm, v))
            self.sd[name + '.module.weight_u'] = u
            self.sd[name + '.module.weight_v'] = v
            sig.append(float(sigma))
        self.calls += 1
        return sig

    # -- building blocks ---------------------------------------------------
    def _bn_fold(self, prefix):
        g = self.sd[prefix + '.weight']
        b = self.sd[prefix + '.bias']
        m = self.sd[prefix + '.running_mean']
        var = self.sd[prefix + '.running_var']
        s = g / torch.sqrt(var + BN_EPS)
        return s, b - m * s

    def _conv_weight(self, name, sigma):
        if name + '.weight' in self.sd:
            return self.sd[name + '.weight']
        return self.sd[name + '.module.weight_bar'] / sigma[name]

    def _block(self, x, prefix, stride, sigma):
        s1, t1 = self._bn_fold(prefix + '.bn1')
        s2, t2 = self._bn_fold(prefix + '.bn2')
        sd, td = self._bn_fold(prefix + '.downsample.1')
        w1 = self._conv_weight(prefix + '.conv1', sigma)
        w2 = self._conv_weight(prefix + '.conv2', sigma)
        wd = self.sd[prefix + '.downsample.0.weight']
        bd = self.sd[prefix + '.downsample.0.bias']
        c = lambda v: v.view(1, -1, 1, 1, 1)
        out = F.conv3d(x, w1, None, stride, 1) * c(s1) + c(t1)
        out = torch.relu(out)
        out = F.conv3d(out, w2, None, 1, 1) * c(s2) + c(t2)
        res = F.conv3d(x, wd, bd, stride, 0) * c(sd) + c(td)
        return torch.relu(out + res)

    @torch.no_grad()
    def forward(self, x, return_intermediates=False):
        """x (B,L,2,H,W) fp32 -> (B,L,20,H,W) fp32.  Advances the SN state by one call."""
        sig = dict(zip(sn_conv_names(), self.sn_step()))
        inter = {}
        x = x.float().permute(0, 2, 1, 3, 4)
        x = F.leaky_relu(F.conv3d(x, self.sd['UNet.head.conv3d.weight'],
                                  self.sd['UNet.head.conv3d.bias'], 1, 1), 0.01)
        inter['head'] = x
        skips = []
        for i in range(4):
            skips.append(x)
            x = self._block(x, f'UNet.encoders.{i}', (1, 2, 2), sig)
            inter[f'enc{i}'] = x
        for i in range(2):
            x = self._block(x, f'UNet.resblocks.{i}', 1, sig)
            inter[f'res{i}'] = x
        for i, skip in enumerate(reversed(skips)):
            hs, ws = skip.shape[3], skip.shape[4]
            hi = (torch.arange(hs) * x.shape[3]) // hs
            wi = (torch.arange(ws) * x.shape[4]) // ws
            up = x[:, :, :, hi][:, :, :, :, wi]
            x = self._block(torch.cat([up, skip], dim=1), f'UNet.decoders.{i}', 1, sig)
            inter[f'dec{i}'] = x
        x = torch.relu(F.conv3d(x, self.sd['UNet.pred.conv3d.weight'], self.sd['UNet.pred.conv3d.bias']))
        out = x.permute(0, 2, 1, 3, 4).contiguous()
        if return_intermediates:
            return out, inter, sig
        return out
