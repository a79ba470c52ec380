"""Row f3 end to end: the data-parallel training loop of gsvc_b200.dp_train on a synthetic video, under torchrun.

    python examples/dp_train.py --iters 250                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \\
        examples/dp_train.py --iters 250                                      # two GPUs: one frame per rank

Targets are frames rendered from a hidden set of free Gaussians; the model starts from a coarse grid of anchors,
trains with the reference's iteration shape (two consecutive frames x front / back view, L1, Adam, densification
every --interval iterations) and must (a) bring the loss down, (b) grow and prune anchors, (c) keep all ranks'
anchors, features, MLP and accumulators bit-identical after every densification round (checked with an all-gather of
checksums).  Prints one summary line per densification round and a final JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from gsvc_b200.dp_train import AnchorModel, DPTrainer
from gsvc_b200.frames import CubeGeometry, synthetic_gaussians
from gsvc_b200.rasterizer import GaussianRasterizationSettings
from gsvc_b200.views import render_toast


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=250)
    ap.add_argument("--interval", type=int, default=5)
    ap.add_argument("--switch", action="store_true", help="sum gradients + statistics with sharding.SwitchAllReduce")
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=192)
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--anchors", type=int, default=6000)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W, H, F, thr = args.width, args.height, args.frames, 0.05
    geom = CubeGeometry(W, H, F)
    bg = torch.zeros(3, device=dev)

    def settings(frame, back):
        fr = geom.frame(frame)
        vm = fr.view_matrix_s if back else fr.view_matrix
        return GaussianRasterizationSettings(image_height=H, image_width=W, x_min=fr.x_min, y_min=fr.y_min, scale=fr.scale,
                                             threshold=thr, bg=bg, scale_modifier=1.0, viewmatrix=vm.permute(1, 0).to(dev),
                                             sh_degree=0, campos=fr.cam_pos, prefiltered=False, debug=False)

    # hidden scene -> target frames (the reference composes a frame as (front + flip(back)) / 2)
    hidden = synthetic_gaussians(20000, geom, 0, F - 1, threshold=thr, seed=11, device=dev)
    targets = {}

    def target(frame):
        if frame not in targets:
            with torch.no_grad():
                img, _, _ = render_toast(settings(frame, False), settings(frame, True), means3D=hidden["means3D"],
                                         opacities=hidden["opacities"], colors_precomp=hidden["colors_precomp"],
                                         scales=hidden["scales"], rotations=hidden["rotations"])
            targets[frame] = img
        return targets[frame]

    g = torch.Generator().manual_seed(3)
    ext = torch.tensor([-2.2 * geom.x_min, -2.2 * geom.y_min, geom.z_of(F - 1) - geom.z_of(0) + 3 * thr])
    org = torch.tensor([1.1 * geom.x_min, 1.1 * geom.y_min, geom.z_of(0) - 1.5 * thr])
    anchors = (org + ext * torch.rand(args.anchors, 3, generator=g)).to(dev)
    # voxel_size 0.4 px: the three growth levels of anchor_growing are 6.4 / 1.6 / 0.4 px voxels (a level only runs if
    # the coarser one added anchors — scene/gaussian_model.py:1374-1377 — so the coarsest must be finer than the
    # initial anchor spacing); Gaussians start 2 px wide
    model = AnchorModel(anchors, n_offsets=4, feat_dim=8, voxel_size=0.4 / geom.scale, seed=5, lr=2e-3,
                        init_scale=2.0 / geom.scale)
    # the K Gaussians of an anchor start spread over its neighbourhood (the reference starts them at the anchor and
    # lets the optimizer push them out over thousands of iterations; growth needs Gaussians that left their voxel)
    model.p["offset"] = (1.5 * torch.randn(args.anchors, 4, 3, generator=g)).to(dev)
    collective = None
    if args.switch and world > 1:
        # the iteration's one exchange through the library's own kernel; room for the anchors to grow 4x
        from gsvc_b200.sharding import SwitchAllReduce
        collective = SwitchAllReduce(4 * sum(n for _, n in model.flat_layout()) // 4 * 4, dev)
    trainer = DPTrainer(model, settings, target, rank=rank, world=world, seed=99, update_interval=args.interval,
                        grad_threshold=2e-5, min_opacity=0.02, collective=collective)
    frame_rng = torch.Generator().manual_seed(17)                   # the same frame draw on every rank (train.py:337)
    first = last = None
    agree, rounds, added, pruned = True, 0, 0, 0
    for it in range(1, args.iters + 1):
        f = int(torch.randint(0, F - 1, (1,), generator=frame_rng))
        rec = trainer.step(f)
        loss = torch.tensor([rec["loss"]], device=dev)
        if world > 1:
            dist.all_reduce(loss)
        if it <= 10:
            first = float(loss) if first is None else first + float(loss)
        if it > args.iters - 10:
            last = float(loss) if last is None else last + float(loss)
        if it % args.interval == 0:
            rounds += 1
            added += rec["added"]
            pruned += rec["pruned"]
            ok = trainer.ranks_agree()
            agree = agree and ok
            if rank == 0 and (rounds % 10 == 0 or not ok):
                print(f"round {rounds:3d} (iteration {it}): anchors {rec['anchors']} (+{rec['added']} -{rec['pruned']}) "
                      f"loss {float(loss):.4f} ranks agree: {ok}", flush=True)
    if rank == 0:
        print(json.dumps({"world": world, "iterations": args.iters, "densification_rounds": rounds, "anchors_start": args.anchors,
                          "anchors_end": model.n_anchors, "added": added, "pruned": pruned,
                          "loss_first10": first / 10, "loss_last10": last / 10, "ranks_agree_every_round": agree,
                          "collective": "torch.distributed all_reduce" if collective is None else
                          f"sharding.SwitchAllReduce ({collective.mode})"}))
    if world > 1:
        dist.destroy_process_group()
    return 0 if agree else 1


if __name__ == "__main__":
    sys.exit(main())
