"""reference models/networks/base_network.py — option holder + the NoiseInjection helpers."""
import torch


class BaseNetwork(torch.nn.Module):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt

    def num_parameters(self):
        return sum(p.numel() for p in self.parameters())

    def print_architecture(self, verbose=False):
        lines = ["-------------------%s---------------------" % type(self).__name__]
        if verbose:
            for name, child in self.named_children():
                lines.append("%s: %3.3fM" % (name, sum(p.numel() for p in child.parameters()) / 1e6))
        lines.append("[Network %s] Total number of parameters : %.3f M" % (type(self).__name__, self.num_parameters() / 1e6))
        print("\n".join(lines))

    def set_requires_grad(self, requires_grad):
        for p in self.parameters():
            p.requires_grad = requires_grad

    def collect_parameters(self, name):
        return [p for m in self.modules() if type(m).__name__ == name for p in m.parameters()]

    def _noise_modules(self):
        return [m for m in self.modules() if type(m).__name__ == "NoiseInjection"]

    def fix_and_gather_noise_parameters(self):
        """Freeze every NoiseInjection to a learnable noise map (reference base_network.py:41-49)."""
        device = next(self.parameters()).device
        params = []
        for m in self._noise_modules():
            assert m.image_size is not None, "One forward call should be made to determine size of noise parameters"
            b, _, h, w = m.image_size
            m.fixed_noise = torch.nn.Parameter(torch.randn(b, 1, h, w, device=device))
            params.append(m.fixed_noise)
        return params

    def remove_noise_parameters(self, name=None):
        for m in self._noise_modules():
            m.fixed_noise = None

    def forward(self, x):
        return x
