"""Dev study (CPU only): centred adjugate plane fit vs the oracle's pivoted-Householder LS on the bench workload.
usage: python tools/dev_fit_study.py [workload] [n_distinct]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, oracle as O
from msf_loam_b200 import synth as S

wl = sys.argv[1] if len(sys.argv) > 1 else "vlp16"
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 3
P = O.default_params()
traj, scans = bench.raw_scans(wl, nd)
mc, ms, queries, _ = bench.build_case(lambda x, r: O.extract_features(P, x, r, None), O.voxel_grid, wl, traj, scans)
rng = np.random.default_rng(5)
tot = 0; worst = 0; nfall = 0; flips = 0
devs = []
for (qc, qs, gt) in queries:
    pose = S.perturb_pose(gt, rng)
    corr, ne, npl, kidx = O.associate_map(P, mc, ms, qc, qs, pose)
    ks = kidx[qc.shape[0]:]
    ok = ks[:, 4] >= 0
    nb = ms[ks[ok]][:, :, :3].astype(np.float64)            # (n,5,3)
    # oracle: LS through the C routine
    n_or = np.zeros((nb.shape[0], 3))
    for i in range(nb.shape[0]):
        x = O.lstsq_5x3(nb[i], -np.ones(5))
        n_or[i] = x / np.linalg.norm(x)
    # fast path: c = mean, S = sum e e^T, n ~ -adj(S) c
    c = ((((nb[:, 0] + nb[:, 1]) + nb[:, 2]) + nb[:, 3]) + nb[:, 4]) / 5.0
    e = nb - c[:, None, :]
    Sm = np.einsum('nki,nkj->nij', e, e)
    a, b, cc, d, ee, f = Sm[:, 0, 0], Sm[:, 0, 1], Sm[:, 0, 2], Sm[:, 1, 1], Sm[:, 1, 2], Sm[:, 2, 2]
    adj = np.stack([np.stack([d * f - ee * ee, cc * ee - b * f, b * ee - cc * d], -1),
                    np.stack([cc * ee - b * f, a * f - cc * cc, b * cc - a * ee], -1),
                    np.stack([b * ee - cc * d, b * cc - a * ee, a * d - b * b], -1)], 1)
    v = -np.einsum('nij,nj->ni', adj, c)
    n_fa = v / np.linalg.norm(v, axis=1, keepdims=True)
    det = a * adj[:, 0, 0] + b * adj[:, 0, 1] + cc * adj[:, 0, 2]
    tr = a + d + f
    ratio = np.linalg.norm(v, axis=1) / (tr ** 2 * np.linalg.norm(c, axis=1))  # 1 / amplification of round-off
    dev = np.linalg.norm(n_fa - n_or, axis=1)
    dd_or = np.abs(np.einsum('nj,nkj->nk', n_or, e)).max(1)
    dd_fa = np.abs(np.einsum('nj,nkj->nk', n_fa, e)).max(1)
    flips += int(((dd_or <= 0.2) != (dd_fa <= 0.2)).sum())
    devs.append(np.stack([dev, ratio, np.linalg.norm(c, axis=1), dd_or], 1))
D = np.concatenate(devs)
print("queries", D.shape[0], "validity flips", flips)
for thr in (1e-12, 1e-11, 1e-10, 1e-9, 1e-8):
    print(f"dev > {thr:g}: {(D[:,0] > thr).sum()}")
o = np.argsort(-D[:, 0])[:15]
print("worst (dev, |v|/(tr^2|c|), |c|, max|dd|):")
for i in o: print("  %.3e  %.3e  %6.2f  %.4f" % tuple(D[i]))
for rthr in (1e-3, 1e-4, 1e-5, 1e-6):
    m = D[:, 1] < rthr
    print(f"ratio < {rthr:g}: {m.sum()} queries ({100*m.mean():.2f} %), max dev among the rest {D[~m,0].max():.3e}")
