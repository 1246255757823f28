import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from emloco_b200.motion_lib import MotionLibSMPL
from emloco_b200.synthetic import synthetic_motion_lib
from oracle import oracle_np as O
arr = synthetic_motion_lib(32, 5)
big = MotionLibSMPL(arr, seed=3)
i2 = big.sample_motions(5000); t2 = big.sample_time(i2)
d2 = big.fetch_amp_obs_demo(5000, motion_ids=i2, motion_times0=t2).cpu().numpy().reshape(5000, 15, 206)
ref = O.amp_obs_demo(arr, i2.cpu().numpy().astype(np.int64), t2.cpu().numpy()).reshape(5000, 15, 206)
bad = np.argwhere(np.abs(d2 - ref) > 1e-4 + 1e-3 * np.abs(ref))
print(len(bad))
ids = i2.cpu().numpy(); tt = t2.cpu().numpy()
seen = set()
for s, k, c in bad[:40]:
    if (s, k) in seen: continue
    seen.add((s, k))
    time = np.float32(tt[s]) + (-np.float32(2/60) * np.float32(k))
    L = arr["motion_lengths"][ids[s]]; nf = arr["motion_num_frames"][ids[s]]; dt = arr["motion_dt"][ids[s]]
    ph = np.clip(np.float32(time) / L, 0, 1)
    print("sample", s, "step", k, "cols", sorted(set(int(x[2]) for x in bad if x[0] == s and x[1] == k)), "time", time, "len", L, "phase*(nf-1)", ph * np.float32(nf - 1),
          "ours", d2[s, k, c], "ref", ref[s, k, c])
