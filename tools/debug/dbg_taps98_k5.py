import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
import numpy as np, torch
import oracle as O
import idsp_b200 as ib
from idsp_b200 import HbfDecCascade, Lanes
from test_gpu_hbf import _oracle_dec_cascade_taps
DEV = "cuda:0"
K = 5
taps = list(ib.hbf_taps_98()[:K])
R, TO = 1 << K, 512 >> K
Ms = [len(taps[K - 1 - s]) for s in range(K)]
offs = np.cumsum([0] + [3 * m - 2 for m in Ms])
print("stage Ms", Ms, "state word offsets", offs.tolist())
for lanes, ntiles in ((8, 1), (8, 2)):
    rng = np.random.default_rng(1)
    n_out = TO * ntiles
    xl = rng.uniform(-1, 1, (lanes, n_out * R)).astype(np.float32)
    cfg = HbfDecCascade(K, taps)
    st = cfg.state(lanes, DEV)
    so = np.zeros(tuple(st.words.shape), np.float32)
    want = np.stack([_oracle_dec_cascade_taps(O, taps, so[:, l], xl[l]) for l in range(lanes)])
    y = torch.empty(n_out * lanes, dtype=torch.float32, device=DEV)
    Lanes(cfg).block(st, torch.from_numpy(xl.reshape(-1)).to(DEV), y, 1)
    got = y.cpu().numpy().reshape(lanes, n_out)
    sg = st.numpy()
    badw = np.nonzero((sg.view(np.uint32) != so.view(np.uint32)).any(1))[0]
    print(f"tiles={ntiles}: output bad {int((got.view(np.uint32) != want.view(np.uint32)).sum())}; bad state words: {badw.tolist()}")
    for s in range(K):
        w = [int(i - offs[s]) for i in badw if offs[s] <= i < offs[s + 1]]
        if w:
            print(f"   stage {s} (M={Ms[s]}, even hist words 0..{Ms[s]-2}, odd hist words {Ms[s]-1}..{3*Ms[s]-3}): bad words {w}")
            for i in w[:3]:
                print("      word", i, "gpu", sg[offs[s] + i, :3], "want", so[offs[s] + i, :3])
