import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package(); O = entry.load_oracle()
K, Kinv = O.reference_K()
n = 2000
for noise in (0.0, 0.1):
    sc = O.synthetic_pair(n, outlier_frac=0.3, noise_px=noise, seed=9)
    x = O.normalise_points(sc["px"], Kinv)
    h = pkg.BatchedPairs(K, Kinv, 1, n, 4096)
    h.set_option(1, 0)
    h.set_points_xy(torch.from_numpy(sc["px"][None]).cuda())
    h.estimate_e(4096, 3, 1e-6); h.refine_e(6); h.pose_candidates(); h.choose_pose(); h.triangulate()
    E0 = h.get_E()[0].astype(np.float64); M0 = h.get_poses()[0][int(h.get_pose_index()[0])].astype(np.float64)
    X0 = h.get_points_host(0).astype(np.float64)
    mask = O.sampson_mask_f32(E0, x, 1e-6)
    act = O.ba_active(x, M0, X0, mask)
    print("noise", noise, "active", act.sum(), "cost with GPU fp32 DLT points", O.ba_cost(x[act], M0[:3,:3], M0[:3,3], X0[:3].T[act]),
          "with fp64 DLT points", O.ba_cost(x[act], M0[:3,:3], M0[:3,3], O.triangulate(x, M0)[:3].T[act]))
    for it in (1, 2, 3, 5, 10):
        h.set_E(E0[None]); 
        # restore pose: re-run pose stages from the same E
        h.pose_candidates(); h.choose_pose(); h.triangulate()
        st = h.bundle_adjust(1, it)[0]
        print("  iters", it, "stats", [float(v) for v in st])
    r = O.bundle_adjust(x, M0, X0, act, 5)
    print("  oracle from the same fp32 points: cost0", r["cost0"], "cost", r["cost"], "accepted", r["accepted"])
    h.close()
