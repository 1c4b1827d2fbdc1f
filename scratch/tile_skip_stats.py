"""CPU analysis (numpy, oracle geometry): how many (128-row tile, offset) steps have at least one rule, and how many
rows of such a step are live, under first-occurrence row order and under Morton row order. Run: python scratch/tile_skip_stats.py [sensor]"""
import sys, numpy as np
sys.path.insert(0, ".")
from mopa_b200 import synth
from oracle import scn_oracle as so

def part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0x1fffff)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1f00000000ffff)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1f0000ff0000ff)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100f00f00f00f00f)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10c30c30c30c30c3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x
def morton(vc):
    return (vc[:, 3].astype(np.uint64) << np.uint64(48)) | (part1by2(vc[:, 0]) << np.uint64(2)) | (part1by2(vc[:, 1]) << np.uint64(1)) | part1by2(vc[:, 2])

def stats(nbr, tile=128):
    K, V = nbr.shape
    nt = (V + tile - 1) // tile
    pad = nt * tile - V
    live = np.pad(nbr >= 0, ((0, 0), (0, pad))).reshape(K, nt, tile)
    cnt = live.sum(2)           # K x nt
    nonempty = cnt > 0
    # 32-row warp granularity: LDGSTS ops with compaction = ceil(live rows in warp/4) (8 lanes per row)
    w = live.reshape(K, nt, tile // 32, 32).sum(3)
    return nonempty.mean(), nonempty.sum(0).mean(), cnt[nonempty].mean(), np.ceil(w / 4).sum() / (K * nt * 4 * 8)

sensor = sys.argv[1] if len(sys.argv) > 1 else "nuscenes"
coords, feats = synth.make_batch(8, sensor, 0)
vc, p2v, _, _ = so.input_layer_rules(coords)
size = 4096
print(f"{sensor}: N={coords.shape[0]}")
vcm = vc[np.argsort(morton(vc), kind='stable')]
for lvl in range(7):
    out = []
    for name, v in (("first", vc), ("morton", vcm)):
        nbr = so.submanifold_rules(v, size)
        f, steps, rows, opsfrac = stats(nbr)
        out.append(f"{name}: nonempty {f:.3f} ({steps:.1f}/27 steps/tile) live rows/step {rows:.1f} compact-op frac {opsfrac:.3f}")
    print(f"L{lvl} V={vc.shape[0]:7d} rules/row={(so.submanifold_rules(vc,size)>=0).sum()/vc.shape[0]:.2f} | " + " | ".join(out))
    if lvl < 6:
        vc, _, _ = so.strided_rules(vc)
        vcm, _, _ = so.strided_rules(vcm)
        size //= 2
