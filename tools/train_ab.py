#!/usr/bin/env python
"""tf32 mode against fp32 mode over a short training run (what does the tensor-core product precision do
to training, not just to one gradient?).

Net3DSeg (UNetSCN + linear head, mopa/models/xmuda_arch.py:82-126) is trained for --steps Adam steps (lr 1e-3, the
reference's optimiser: configs/nuscenes/*/xmuda_pl_mopa.yaml) on synthetic nuScenes-shaped scans with a label the network
can learn from occupancy alone: the number of occupied voxels among a point's 26 neighbours (0, 1, 2, 3, 4 or more). Same initial
weights, same batches, same order in both modes; everything else (BatchNorm, optimiser, loss) is identical fp32 code.
Prints the loss every --every steps, the held-out accuracy at the end, and the relative distance between the two runs'
final weights.  python tools/train_ab.py [--steps 200] > profiles/r02_train_ab.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def density_labels(coords):
    """5 classes from the count of occupied voxels among the 26 neighbours of a point's voxel (0, 1, 2, 3, 4+)."""
    c = coords.astype(np.int64)
    key = (c[:, 3] << 48) | (c[:, 0] << 32) | (c[:, 1] << 16) | c[:, 2]
    uniq = np.unique(key)
    cnt = np.zeros(c.shape[0], np.int64)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                if dx == dy == dz == 0:
                    continue
                x, y, z = c[:, 0] + dx, c[:, 1] + dy, c[:, 2] + dz
                ok = (x >= 0) & (y >= 0) & (z >= 0)
                k = (c[:, 3] << 48) | (x << 32) | (y << 16) | z
                cnt += ok & np.isin(k, uniq)
    return np.minimum(cnt, 4)  # 0, 1, 2, 3, >= 4 occupied neighbours: roughly 24 / 16 / 19 / 17 / 24 % of the points


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--every", type=int, default=20)
    ap.add_argument("--batches", type=int, default=16)
    ap.add_argument("--scans", type=int, default=2)
    a = ap.parse_args()
    import mopa_b200.scn as scn
    from mopa_b200 import synth
    from mopa_b200.unet_scn import Net3DSeg

    data = []
    for s in range(a.batches + 2):  # the last two are held out
        coords, feats = synth.make_batch(a.scans, "nuscenes", 100 + s)
        data.append((torch.from_numpy(coords), torch.from_numpy(feats).cuda(), torch.from_numpy(density_labels(coords)).cuda()))
    train, held = data[:a.batches], data[a.batches:]
    torch.manual_seed(0)
    init = Net3DSeg(5, dual_head=False, backbone_3d="SCN", backbone_3d_kwargs={"in_channels": 1}).cuda().state_dict()
    init = {k: v.clone() for k, v in init.items()}

    curves, finals, accs = {}, {}, {}
    for mode in ("fp32", "tf32", "fp32 again"):  # the rerun calibrates: how far do two IDENTICAL runs drift (atomics order)?
        scn.set_precision(mode.split()[0])
        model = Net3DSeg(5, dual_head=False, backbone_3d="SCN", backbone_3d_kwargs={"in_channels": 1}).cuda()
        model.load_state_dict(init)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        model.train()
        losses = []
        for step in range(a.steps):
            c, f, y = train[step % len(train)]
            opt.zero_grad(set_to_none=True)
            loss = F.cross_entropy(model({"x": [c, f]})["seg_logit"], y)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        model.eval()
        hit = tot = 0
        with torch.no_grad():
            for c, f, y in held:
                p = model({"x": [c, f]})["seg_logit"].argmax(1)
                hit += int((p == y).sum())
                tot += y.numel()
        curves[mode], accs[mode] = losses, hit / tot
        finals[mode] = {k: v.detach().double().cpu().clone() for k, v in model.named_parameters()}
        finals[mode + "/delta"] = {k: finals[mode][k] - init[k].double().cpu() for k in finals[mode]}

    print("# tools/train_ab.py: Net3DSeg (UNetSCN + linear head), %d Adam steps (lr 1e-3), %d scans/step of ~32.9k points, 5-class"
          % (a.steps, a.scans))
    print("# neighbour-density labels; same initial weights and batches in both modes. fp32 = 3xTF32 split products on mma.sync,")
    print("# tf32 = single TF32 products on tcgen05 (the default training mode).")
    print("%6s %12s %12s %10s" % ("step", "loss fp32", "loss tf32", "tf32-fp32"))
    for s in list(range(0, a.steps, a.every)) + [a.steps - 1]:
        lo, hi = max(0, s - 4), min(a.steps, s + 5)  # 9-step window: single-step losses jump with the batch
        m32, mtf = np.mean(curves["fp32"][lo:hi]), np.mean(curves["tf32"][lo:hi])
        print("%6d %12.5f %12.5f %+10.5f" % (s, m32, mtf, mtf - m32))
    print("mean loss of the last 20 steps: fp32 %.5f  tf32 %.5f" % (np.mean(curves["fp32"][-20:]), np.mean(curves["tf32"][-20:])))
    print("held-out accuracy (2 batches): fp32 %.4f  tf32 %.4f" % (accs["fp32"], accs["tf32"]))

    def flat(d):
        return torch.cat([d[k].flatten() for k in sorted(d)])
    w32, wtf = flat(finals["fp32"]), flat(finals["tf32"])
    d32, dtf = flat(finals["fp32/delta"]), flat(finals["tf32/delta"])
    print("final parameters: relative L2 distance %.3e, cosine %.6f" % (float((wtf - w32).norm() / w32.norm()),
                                                                       float(torch.dot(wtf, w32) / (wtf.norm() * w32.norm()))))
    print("parameter UPDATES (final - initial): relative L2 distance %.3e, cosine %.6f; |update| / |initial| = %.3f" % (
        float((dtf - d32).norm() / d32.norm()), float(torch.dot(dtf, d32) / (dtf.norm() * d32.norm())),
        float(d32.norm() / flat({k: init[k].double().cpu() for k in finals["fp32"]}).norm())))
    wre, dre = flat(finals["fp32 again"]), flat(finals["fp32 again/delta"])
    print("the same for fp32 against its own rerun: parameters %.3e, updates %.3e (cosine %.6f); loss of the last 20 steps %.5f"
          % (float((wre - w32).norm() / w32.norm()), float((dre - d32).norm() / d32.norm()),
             float(torch.dot(dre, d32) / (dre.norm() * d32.norm())), np.mean(curves["fp32 again"][-20:])))
    print("(Adam divides every gradient by its running magnitude: where the gradient is noise-sized, rounding differences decide")
    print(" the direction of a full-sized step, so two runs drift apart in weight space while following the same loss curve.)")

if __name__ == "__main__":
    main()
