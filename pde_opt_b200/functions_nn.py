"""Neural closures for mu (and D): torch mirrors of the reference's equinox modules
pde_opt/numerics/functions/cnn.py:13-102 (PeriodicConvBlock, PeriodicCNN) and mixer_mlp.py:13-86 (MixerBlock,
Mixer2d) — same constructor arguments, same layer structure (circular 'SAME' convolutions + GELU; patch embedding,
token / channel mixing MLPs with LayerNorm, transposed-convolution read-out).

They are closures of the WHOLE field, not pointwise families, so they cannot be folded into the fused stepper.  The
equation evaluates them on the whole batch ([B, nx, ny] in, [B, nx, ny] out; torch / cuDNN, i.e. library code outside
the stepping kernels) and hands the result to pdeopt_rhs_given_mu_batched (CUDA stencils) and
pdeopt_sifs_filter_batched (the fused spectral filter): `SemiImplicitFourierSpectral.step / rollout` take that route
for any equation whose closures are not enumerated.  Weights are initialised by torch (`key` seeds a torch generator;
jax's threefry streams cannot be reproduced); load reference weights with `load_state_dict` for bit-level comparisons."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _seed(key):
    if key is None:
        return None
    return torch.Generator().manual_seed(int(key))


def _reinit(module, gen):
    """Deterministic initialisation from `gen` (uniform +-1/sqrt(fan_in), the equinox / torch default family)."""
    if gen is None:
        return
    with torch.no_grad():
        for p in module.parameters():
            fan_in = p[0].numel() if p.dim() > 1 else p.numel()
            bound = 1.0 / max(1.0, float(fan_in)) ** 0.5
            p.copy_((torch.rand(p.shape, generator=gen) * 2 - 1) * bound)


class PeriodicConvBlock(nn.Module):
    """Conv2d -> activation with periodic padding (cnn.py:13-43)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, act=F.gelu, *, key=None):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=1, padding=kernel_size // 2, padding_mode="circular", bias=True)
        self.act = act
        _reinit(self.conv, _seed(key))

    def forward(self, x):
        return self.act(self.conv(x))


class PeriodicCNN(nn.Module):
    """Stack of periodic conv blocks; the final convolution has no activation (cnn.py:46-102).  Called on a field
    [nx, ny] (as in the reference, one channel) or a batch [B, nx, ny]; returns the same shape."""

    is_field_closure = True

    def __init__(self, in_channels=1, hidden_channels=(32, 64, 64), out_channels=None, kernel_size=3, act=F.gelu, *, key=None):
        super().__init__()
        assert kernel_size % 2 == 1, "Use odd kernels to avoid off-by-one alignment."
        out_channels = in_channels if out_channels is None else out_channels
        blocks, c_prev = [], in_channels
        for c_next in hidden_channels:
            blocks.append(PeriodicConvBlock(c_prev, c_next, kernel_size, act))
            c_prev = c_next
        blocks.append(nn.Conv2d(c_prev, out_channels, kernel_size, stride=1, padding=kernel_size // 2, padding_mode="circular", bias=True))
        self.layers = nn.ModuleList(blocks)
        _reinit(self, _seed(key))

    def forward(self, x):
        single = x.dim() == 2
        y = (x[None] if single else x)[:, None]  # [B, 1, nx, ny]
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):  # float32 convolutions: parity, not TF32 speed
            for layer in self.layers:
                y = layer(y)
        y = y[:, 0]
        return y[0] if single else y


class MixerBlock(nn.Module):
    """mixer_mlp.py:13-37: token mixing over patches, channel mixing over the hidden size, pre-LayerNorm, residual."""

    def __init__(self, num_patches, hidden_size, mix_patch_size, mix_hidden_size):
        super().__init__()
        self.patch_mixer = nn.Sequential(nn.Linear(num_patches, mix_patch_size), nn.ReLU(), nn.Linear(mix_patch_size, num_patches))
        self.hidden_mixer = nn.Sequential(nn.Linear(hidden_size, mix_hidden_size), nn.ReLU(), nn.Linear(mix_hidden_size, hidden_size))
        self.norm1 = nn.LayerNorm((hidden_size, num_patches))
        self.norm2 = nn.LayerNorm((num_patches, hidden_size))

    def forward(self, y):  # [B, c, p]
        y = y + self.patch_mixer(self.norm1(y))
        y = y.transpose(1, 2)
        y = y + self.hidden_mixer(self.norm2(y))
        return y.transpose(1, 2)


class Mixer2d(nn.Module):
    """mixer_mlp.py:40-86 (eqx.nn.MLP's default activation is relu)."""

    is_field_closure = True

    def __init__(self, img_size, patch_size, hidden_size, mix_patch_size, mix_hidden_size, num_blocks, *, key=None):
        super().__init__()
        input_size, height, width = img_size
        assert height % patch_size == 0 and width % patch_size == 0
        self.grid = (height // patch_size, width // patch_size)
        num_patches = self.grid[0] * self.grid[1]
        self.conv_in = nn.Conv2d(input_size, hidden_size, patch_size, stride=patch_size)
        self.conv_out = nn.ConvTranspose2d(hidden_size, input_size, patch_size, stride=patch_size)
        self.blocks = nn.ModuleList([MixerBlock(num_patches, hidden_size, mix_patch_size, mix_hidden_size) for _ in range(num_blocks)])
        self.norm = nn.LayerNorm((hidden_size, num_patches))
        _reinit(self, _seed(key))

    def forward(self, x):
        single = x.dim() == 2
        y = (x[None] if single else x)[:, None]
        y = self.conv_in(y)
        B, c = y.shape[:2]
        y = y.reshape(B, c, -1)
        for blk in self.blocks:
            y = blk(y)
        y = self.norm(y).reshape(B, c, *self.grid)
        y = self.conv_out(y)[:, 0]
        return y[0] if single else y
