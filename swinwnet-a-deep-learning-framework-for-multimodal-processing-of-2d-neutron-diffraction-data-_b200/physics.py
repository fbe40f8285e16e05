"""GPU front end of the physics metrics: the reference's ``Qwrapper`` (Diffraction_metrics.py:9-70) with the same
constructor arguments and the same ``tensor_to_d`` return value, but the whole batch is reduced by one kernel launch
(``swn_dspace_histogram``) and copied to the host once.  Peak finding / comparison stay with the caller (scipy), as in
the reference."""
import numpy as np
import torch

from . import ops


class Qwrapper:
    def __init__(self, theta_range=(-170, 170), L_range=(0.1, 10), fixed_centers=None, device="cuda"):
        if fixed_centers is None:
            raise ValueError("fixed_centers must be provided")
        self.theta_range, self.L_range = theta_range, L_range
        self.device = torch.device(device)
        centers = torch.as_tensor(np.asarray(fixed_centers), dtype=torch.float32)
        self.centers = centers.to(self.device)
        edges = torch.zeros(len(centers) + 1, dtype=torch.float32)                 # Diffraction_metrics.py:25-31
        edges[1:-1] = (centers[:-1] + centers[1:]) * 0.5
        edges[0] = centers[0] - (centers[1] - centers[0]) * 0.5
        edges[-1] = centers[-1] + (centers[-1] - centers[-2]) * 0.5
        self.edges = edges.to(self.device)
        self._maps = {}

    def bin_map(self, H, W):
        """int32 [H*W] pixel -> bin index (-1: d > 7.5), built once per geometry with the reference's own fp32 ops (:43-62)
        on the CPU, so that pixels sitting on a bin edge fall where the CPU reference puts them (a device sin differs in
        the last ulp)."""
        key = (H, W)
        if key not in self._maps:
            theta = torch.deg2rad(torch.linspace(*self.theta_range, W))
            lam = torch.linspace(*self.L_range, H)
            L_grid, theta_grid = torch.meshgrid(lam, theta, indexing="ij")
            d = L_grid / (2 * torch.sin(torch.abs(theta_grid) * 0.5))
            idx = (torch.bucketize(d, self.edges.cpu()) - 1).clamp(0, len(self.centers) - 1)
            idx = torch.where(d <= 7.5, idx, torch.full_like(idx, -1))
            self._maps[key] = idx.to(torch.int32).reshape(-1).contiguous().to(self.device)
        return self._maps[key]

    def tensor_to_d_batched(self, batch_tensor):
        """[B, C, H, W] -> I(d) as one [B, n_bins] fp32 CUDA tensor."""
        if batch_tensor.dim() != 4:
            raise ValueError("Expected tensor [B,1,H,W]")
        x = batch_tensor.to(self.device)
        return ops.dspace_histogram(x, self.bin_map(x.shape[2], x.shape[3]), len(self.centers))

    def tensor_to_d(self, batch_tensor):
        """reference return format: list of {"d": centers, "I": summed intensities} numpy dicts, one per sample."""
        I = self.tensor_to_d_batched(batch_tensor).cpu().numpy()
        d = self.centers.detach().cpu().numpy()
        return [{"d": d, "I": I[b]} for b in range(I.shape[0])]
