"""Data-parallel training loop at the rasterizer + anchor level — SURVEY.md §8f row f3.

Mirrors the hot loop of /root/reference/pipeline/train.py:325-581 for everything that touches this path, one
process per GPU:

  per iteration (train.py:337-387)   two consecutive frames x (front, back view) = 4 rasterizer views.  The views are
                                     dealt round-robin to the ranks (1 GPU: all four in turn; 2 GPUs: one frame each;
                                     4 GPUs: one view each); each view runs the reference's render() chain on this
                                     path's own pieces: visible_filter_compact (prefilter_voxel) -> gather -> per-anchor
                                     MLP -> fused epilogue (generate.py) -> rasterizer -> L1 against the target.
  training_statis x 4 (:559-562)     per view, into DELTA buffers: opacity_accum / anchor_demon per anchor,
                                     offset_gradient_accum / offset_denom per (anchor, offset)
                                     (scene/gaussian_model.py:1298-1314).
  ONE all-reduce                     a flat fp32 buffer [all parameter gradients | the four statistic deltas]: gradients
                                     are averaged over the views of the iteration, statistics are summed — so every rank
                                     holds the statistic of ALL four views, exactly what the single-GPU loop accumulates.
  adjust_anchor (:564-566)           grow (anchor_growing, scene/gaussian_model.py:1362-1448: three levels of voxel
                                     candidates above the gradient threshold, a random thinning with torch.rand_like
                                     at :1369, de-duplication against existing anchors) and prune (:1451-1505), with
                                     the random thinning drawn from a generator seeded identically on every rank, so
                                     all ranks grow and prune THE SAME anchors and never need to exchange them.
  optimizer step (:577-579)          Adam on identical gradients -> identical parameters on every rank.

What is NOT here (out of scope): the hash-grid / entropy models, the codec, SSIM, the optical-flow term,
learning-rate schedules, checkpointing.  The per-anchor "MLP" is one linear layer + activations — a stand-in with the
reference's interface (opacity [K], colour [3K], scale/rotation [7K], offset [3K] per visible anchor), because the
generator's networks are not on the rasterizer path; the DATA FLOW around them is the reference's.
There is no CPU fallback for the render chain; `adjust_anchor`, the statistics and the all-reduce layout are plain
tensor code and are also exercised on CPU with gloo (tests/test_host.py).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

PARAM_NAMES = ("anchor", "offset", "scaling", "rotation", "anchor_feat", "mlp_w", "mlp_b")
ANCHOR_PARAMS = ("anchor", "offset", "scaling", "rotation", "anchor_feat")      # first dimension = number of anchors


class AnchorModel:
    """The anchor-level state of scene/gaussian_model.py::GaussianModel that this path reads and densifies:
    _anchor [N,3], _offset [N,K,3], _scaling [N,6] (log), _rotation [N,4], _anchor_feat [N,F], a stub MLP, the four
    densification accumulators, and Adam moments that follow the anchors through grow / prune
    (cat_tensors_to_optimizer / prune_anchor)."""

    def __init__(self, anchor: torch.Tensor, n_offsets: int = 4, feat_dim: int = 8, voxel_size: float = 0.01,
                 seed: int = 0, lr: float = 2e-3, init_scale: Optional[float] = None):
        dev = anchor.device
        g = torch.Generator().manual_seed(seed)
        N, K, F = int(anchor.shape[0]), int(n_offsets), int(feat_dim)
        self.K, self.F, self.voxel_size, self.lr = K, F, float(voxel_size), float(lr)
        self.update_depth, self.update_init_factor, self.update_hierachy_factor = 3, 16, 4
        self.p: Dict[str, torch.Tensor] = {
            "anchor": anchor.detach().clone().float(),
            "offset": torch.zeros(N, K, 3, device=dev),
            "scaling": torch.full((N, 6), math.log(voxel_size if init_scale is None else init_scale), device=dev),
            "rotation": torch.tensor([1.0, 0.0, 0.0, 0.0], device=dev).repeat(N, 1),
            "anchor_feat": (0.5 * torch.randn(N, F, generator=g)).to(dev),
            "mlp_w": (0.5 * torch.randn(F + 1, 14 * K, generator=g) / math.sqrt(F + 1)).to(dev),
            "mlp_b": torch.zeros(14 * K, device=dev),
        }
        self.adam = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in self.p.items()}
        self.step_no = 0
        self.x_bound_min = torch.full((1, 3), -10.0, device=dev)
        self.x_bound_max = torch.full((1, 3), 10.0, device=dev)
        self.reset_statistics()

    # ---- accessors named after the reference's properties
    @property
    def n_anchors(self) -> int:
        return int(self.p["anchor"].shape[0])

    @property
    def device(self):
        return self.p["anchor"].device

    def get_scaling(self, scaling=None):
        return torch.exp(self.p["scaling"] if scaling is None else scaling)       # scaling_activation

    def reset_statistics(self):
        N, K, dev = self.n_anchors, self.K, self.device
        self.opacity_accum = torch.zeros(N, 1, device=dev)
        self.anchor_demon = torch.zeros(N, 1, device=dev)
        self.offset_gradient_accum = torch.zeros(N * K, 1, device=dev)
        self.offset_denom = torch.zeros(N * K, 1, device=dev)

    # ---- flat all-reduce buffer: [gradients of every parameter | the four statistic deltas]
    def flat_layout(self) -> List[Tuple[str, int]]:
        N, K = self.n_anchors, self.K
        return [(k, int(self.p[k].numel())) for k in PARAM_NAMES] + \
               [("opacity_accum", N), ("anchor_demon", N), ("offset_gradient_accum", N * K), ("offset_denom", N * K)]

    def state_hash(self) -> torch.Tensor:
        """A few floats that differ if any rank's anchors differ (count, and position / feature checksums)."""
        a = self.p["anchor"].double()
        return torch.stack([torch.tensor(float(self.n_anchors), device=self.device, dtype=torch.float64), a.sum(),
                            (a * a).sum(), self.p["anchor_feat"].double().sum(), self.p["mlp_w"].double().sum(),
                            self.offset_denom.double().sum(), self.anchor_demon.double().sum()])

    # ---- Adam (identical gradients on every rank -> identical parameters)
    def optimizer_step(self, grads: Dict[str, torch.Tensor], b1=0.9, b2=0.999, eps=1e-15):
        self.step_no += 1
        c1, c2 = 1 - b1 ** self.step_no, 1 - b2 ** self.step_no
        for k, g in grads.items():
            m, v = self.adam[k]
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            self.p[k].addcdiv_(m / c1, (v / c2).sqrt_().add_(eps), value=-self.lr)

    # ---- training_statis (scene/gaussian_model.py:1298-1314) into delta buffers
    @staticmethod
    def statistics_of_view(deltas, N, K, visible_idx, neural_opacity, selection_mask, radii, means2D_grad):
        """`visible_idx` [n_vis] anchors of this view; `neural_opacity` [n_vis*K,1]; `selection_mask` [n_vis*K] bool
        (generated Gaussians); `radii` [M] of the generated Gaussians; `means2D_grad` [M,3]."""
        d_op, d_dem, d_acc, d_den = deltas
        idx = visible_idx.long()
        temp = neural_opacity.detach().view(-1).clamp_min(0).view(-1, K)
        d_op.index_add_(0, idx, temp.sum(dim=1, keepdim=True))
        d_dem.index_add_(0, idx, torch.ones(idx.numel(), 1, device=d_dem.device))
        # rows of the [N*K] accumulators that belong to generated Gaussians which were actually drawn
        off_rows = (idx.view(-1, 1) * K + torch.arange(K, device=idx.device).view(1, -1)).view(-1)[selection_mask]
        drawn = radii > 0
        rows = off_rows[drawn]
        norm = torch.norm(means2D_grad[drawn, :2], dim=-1, keepdim=True)
        d_acc.index_add_(0, rows, norm)
        d_den.index_add_(0, rows, torch.ones_like(norm))

    # ---- adjust_anchor (scene/gaussian_model.py:1362-1505)
    def adjust_anchor(self, generator: torch.Generator, check_interval=100, success_threshold=0.8,
                      grad_threshold=0.0002, min_opacity=0.005) -> Tuple[int, int]:
        """Grow then prune.  `generator`: a CPU torch.Generator every rank seeded with the same value (the reference
        draws torch.rand_like on the device, :1369; a device draw is reproducible across ranks too, but a host draw
        does not depend on the GPU model).  Returns (anchors added, anchors pruned)."""
        K, dev = self.K, self.device
        grads = self.offset_gradient_accum / self.offset_denom
        grads[grads.isnan()] = 0.0
        grads_norm = torch.norm(grads, dim=-1)
        offset_mask = (self.offset_denom > check_interval * success_threshold * 0.5).squeeze(dim=1)
        n_before = self.n_anchors
        self._anchor_growing(grads_norm, grad_threshold, offset_mask, generator)
        added = self.n_anchors - n_before
        self.offset_denom[offset_mask] = 0
        self.offset_gradient_accum[offset_mask] = 0
        pad = self.n_anchors * K - self.offset_denom.shape[0]
        self.offset_denom = torch.cat([self.offset_denom, torch.zeros(pad, 1, device=dev)], dim=0)
        self.offset_gradient_accum = torch.cat([self.offset_gradient_accum, torch.zeros(pad, 1, device=dev)], dim=0)
        prune_mask = (self.opacity_accum < min_opacity * self.anchor_demon).squeeze(dim=1)
        anchors_mask = (self.anchor_demon > check_interval * success_threshold).squeeze(dim=1)
        prune_mask = torch.logical_and(prune_mask, anchors_mask)
        keep = ~prune_mask
        self.offset_denom = self.offset_denom.view(-1, K)[keep].reshape(-1, 1)
        self.offset_gradient_accum = self.offset_gradient_accum.view(-1, K)[keep].reshape(-1, 1)
        self.opacity_accum[anchors_mask] = 0
        self.anchor_demon[anchors_mask] = 0
        self.opacity_accum = self.opacity_accum[keep]
        self.anchor_demon = self.anchor_demon[keep]
        pruned = int(prune_mask.sum())
        if pruned:
            for k in ANCHOR_PARAMS:
                self.p[k] = self.p[k][keep].contiguous()
                self.adam[k] = tuple(t[keep].contiguous() for t in self.adam[k])
        return added, pruned

    def _anchor_growing(self, grads, threshold, offset_mask, generator):
        K, dev = self.K, self.device
        init_length = self.n_anchors * K
        for i in range(self.update_depth):
            cur_threshold = threshold * ((self.update_hierachy_factor // 2) ** i)
            candidate_mask = torch.logical_and(grads >= cur_threshold, offset_mask)
            rand = torch.rand(candidate_mask.shape[0], generator=generator).to(dev)      # rand_like, :1369
            candidate_mask = torch.logical_and(candidate_mask, rand > (0.5 ** (i + 1)))
            length_inc = self.n_anchors * K - init_length
            if length_inc == 0:
                if i > 0:
                    continue
            else:
                candidate_mask = torch.cat([candidate_mask, torch.zeros(length_inc, dtype=torch.bool, device=dev)])
            anchor, scaling = self.p["anchor"], self.get_scaling()
            all_xyz = anchor.unsqueeze(1) + self.p["offset"] * scaling[:, :3].unsqueeze(1)
            size_factor = self.update_init_factor // (self.update_hierachy_factor ** i)
            cur_size = self.voxel_size * size_factor
            grid_coords = torch.round(anchor / cur_size).int()
            selected_xyz = all_xyz.view(-1, 3)[candidate_mask]
            if selected_xyz.shape[0] == 0:
                continue
            selected_grid = torch.round(selected_xyz / cur_size).int()
            uniq, inverse = torch.unique(selected_grid, return_inverse=True, dim=0)
            # drop candidates whose voxel already holds an anchor: a 3-column integer join through one int64 key
            key = lambda c: (c[:, 0].long() + (1 << 20)) * (1 << 42) + (c[:, 1].long() + (1 << 20)) * (1 << 21) + (c[:, 2].long() + (1 << 20))
            exists = torch.isin(key(uniq), key(grid_coords))
            new_anchor = uniq[~exists].float() * cur_size
            M = int(new_anchor.shape[0])
            if M == 0:
                continue
            feat_rep = self.p["anchor_feat"].unsqueeze(1).expand(-1, K, -1).reshape(-1, self.F)[candidate_mask]
            new_feat = torch.full((uniq.shape[0], self.F), -float("inf"), device=dev)
            new_feat.scatter_reduce_(0, inverse.view(-1, 1).expand(-1, self.F), feat_rep, reduce="amax")   # scatter_max
            new_feat = new_feat[~exists]
            new = {"anchor": new_anchor, "offset": torch.zeros(M, K, 3, device=dev),
                   "scaling": torch.full((M, 6), math.log(cur_size), device=dev),
                   "rotation": torch.tensor([1.0, 0.0, 0.0, 0.0], device=dev).repeat(M, 1), "anchor_feat": new_feat}
            for k, t in new.items():                       # cat_tensors_to_optimizer: new rows start with zero moments
                self.p[k] = torch.cat([self.p[k], t], dim=0)
                self.adam[k] = tuple(torch.cat([s, torch.zeros_like(t)], dim=0) for s in self.adam[k])
            self.anchor_demon = torch.cat([self.anchor_demon, torch.zeros(M, 1, device=dev)], dim=0)
            self.opacity_accum = torch.cat([self.opacity_accum, torch.zeros(M, 1, device=dev)], dim=0)


def views_of_iteration(frame_idx: int) -> List[Tuple[int, bool]]:
    """train.py:337-387: (frame, back?) of the four render() calls of one iteration."""
    return [(frame_idx, False), (frame_idx, True), (frame_idx + 1, False), (frame_idx + 1, True)]


def views_for_rank(views: Sequence, rank: int, world: int) -> List:
    return [v for i, v in enumerate(views) if i % world == rank]


def allreduce_iteration(model: AnchorModel, grads: Dict[str, torch.Tensor], deltas, n_views_total: int,
                        world: int, collective=None) -> Dict[str, torch.Tensor]:
    """ONE collective per iteration: [parameter gradients | statistic deltas] summed over the ranks; gradients are then
    divided by the number of views (the reference sums the four views' losses; the mean keeps the step size
    independent of how many views an iteration has), statistics stay sums and are added into the accumulators."""
    layout = model.flat_layout()
    flat = torch.cat([grads[k].reshape(-1) for k in PARAM_NAMES] + [d.reshape(-1) for d in deltas])
    if world > 1:
        if collective is not None:
            collective.sum_(flat)          # sharding.SwitchAllReduce: this library's own exchange kernel (NVLink / NVSwitch)
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    out, off = {}, 0
    for name, n in layout:
        out[name] = flat[off:off + n]
        off += n
    g = {k: (out[k] / n_views_total).view_as(model.p[k]) for k in PARAM_NAMES}
    model.opacity_accum += out["opacity_accum"].view(-1, 1)
    model.anchor_demon += out["anchor_demon"].view(-1, 1)
    model.offset_gradient_accum += out["offset_gradient_accum"].view(-1, 1)
    model.offset_denom += out["offset_denom"].view(-1, 1)
    return g


class DPTrainer:
    """One process per GPU.  `targets(frame) -> [3,H,W]`; `settings(frame, back)` -> GaussianRasterizationSettings."""

    def __init__(self, model: AnchorModel, settings_fn, target_fn, rank: int = 0, world: int = 1, seed: int = 1234,
                 update_interval: int = 100, densify_from: int = 0, grad_threshold: float = 0.0002,
                 min_opacity: float = 0.005, success_threshold: float = 0.8, collective=None):
        self.m, self.settings_fn, self.target_fn = model, settings_fn, target_fn
        self.collective = collective        # None: torch.distributed all_reduce (NCCL / gloo)
        self.rank, self.world, self.seed = rank, world, seed
        self.update_interval, self.densify_from = update_interval, densify_from
        self.grad_threshold, self.min_opacity, self.success_threshold = grad_threshold, min_opacity, success_threshold
        self.iteration = 0
        self.log: List[dict] = []

    def render_view(self, frame: int, back: bool, leaves: Dict[str, torch.Tensor]):
        """The reference's render() on this path's pieces (renderer.py:14-119)."""
        from .generate import neural_gaussians_epilogue
        from .rasterizer import GaussianRasterizer
        m, K = self.m, self.m.K
        rs = self.settings_fn(frame, back)
        rast = GaussianRasterizer(raster_settings=rs)
        scaling = m.get_scaling(leaves["scaling"])
        rot = torch.nn.functional.normalize(leaves["rotation"])
        # prefilter_voxel (preprocess.py:30-118) fused with the compaction of the visible anchors
        idx, _ = rast.visible_filter_compact(means3D=leaves["anchor"].detach(), scales=scaling.detach()[:, :3],
                                             rotations=rot.detach(), want_radii=False)
        feat = leaves["anchor_feat"].index_select(0, idx.long())
        z = (leaves["anchor"].index_select(0, idx.long())[:, 2:] - rs.campos[2].to(feat.device))     # ob_view, :225-228
        h = torch.cat([feat, z], dim=1) @ leaves["mlp_w"] + leaves["mlp_b"]                          # the stub MLP
        nop, col = torch.tanh(h[:, :K]), torch.sigmoid(h[:, K:4 * K])
        sr, noff = h[:, 4 * K:11 * K], 0.1 * torch.tanh(h[:, 11 * K:14 * K])
        masks = torch.ones(m.n_anchors, K, 1, device=feat.device)
        g = neural_gaussians_epilogue(leaves["anchor"], leaves["offset"], scaling, masks, idx, nop, col, sr, noff,
                                      m.x_bound_min, m.x_bound_max)
        means2D = torch.zeros_like(g.xyz, requires_grad=True)
        image, radii, n = rast(means3D=g.xyz, means2D=means2D, shs=None, colors_precomp=g.color, opacities=g.opacity,
                               scales=g.scaling, rotations=g.rot, cov3D_precomp=None)
        return image, radii, means2D, idx, g

    def step(self, frame_idx: int) -> dict:
        m, K = self.m, self.m.K
        self.iteration += 1
        views = views_of_iteration(frame_idx)
        mine = views_for_rank(views, self.rank, self.world)
        leaves = {k: v.detach().requires_grad_(True) for k, v in m.p.items()}
        N, dev = m.n_anchors, m.device
        deltas = [torch.zeros(N, 1, device=dev), torch.zeros(N, 1, device=dev), torch.zeros(N * K, 1, device=dev),
                  torch.zeros(N * K, 1, device=dev)]
        loss_total, per_view = 0.0, []
        for frame, back in mine:
            image, radii, means2D, idx, g = self.render_view(frame, back, leaves)
            target = self.target_fn(frame)
            img = torch.flip(image, dims=(-1,)) if back else image        # the back view is the x-mirror (train.py:370)
            loss = (img - target).abs().mean()
            loss_total = loss_total + loss
            per_view.append((radii, means2D, idx, g))
        # one backward for the parameter gradients AND the screen-space gradients the statistic needs (means2D.grad)
        grads = {k: torch.zeros_like(m.p[k]) for k in PARAM_NAMES}
        if mine:
            got = torch.autograd.grad(loss_total, [leaves[k] for k in PARAM_NAMES] + [pv[1] for pv in per_view],
                                      allow_unused=True)
            for k, gr in zip(PARAM_NAMES, got[:len(PARAM_NAMES)]):
                if gr is not None:
                    grads[k] = gr
            with torch.no_grad():
                for (radii, means2D, idx, g), m2g in zip(per_view, got[len(PARAM_NAMES):]):
                    m2g = torch.zeros_like(means2D) if m2g is None else m2g
                    AnchorModel.statistics_of_view(deltas, N, K, idx, g.neural_opacity, g.mask, radii, m2g)
        with torch.no_grad():
            g_avg = allreduce_iteration(m, grads, deltas, len(views), self.world, self.collective)
            added = pruned = 0
            if self.iteration > self.densify_from and self.iteration % self.update_interval == 0:
                gen = torch.Generator().manual_seed(self.seed + self.iteration)     # the same on every rank
                # anchors change: this iteration's gradients are applied first (shapes still match)
                m.optimizer_step(g_avg)
                added, pruned = m.adjust_anchor(gen, check_interval=self.update_interval,
                                                success_threshold=self.success_threshold,
                                                grad_threshold=self.grad_threshold, min_opacity=self.min_opacity)
            else:
                m.optimizer_step(g_avg)
        rec = {"iteration": self.iteration, "loss": float(loss_total) if mine else 0.0, "anchors": m.n_anchors,
               "added": added, "pruned": pruned}
        self.log.append(rec)
        return rec

    def ranks_agree(self) -> bool:
        """All ranks hold the same anchors, features, MLP and accumulators (checksums compared over the group)."""
        h = self.m.state_hash()
        if self.world == 1:
            return True
        hs = [torch.zeros_like(h) for _ in range(self.world)]
        dist.all_gather(hs, h)
        return all(torch.equal(hs[0], x) for x in hs[1:])
