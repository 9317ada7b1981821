#!/usr/bin/env python
"""Tuning aid (GPU): phase timeline of the onesweep passes.  python scripts/sort_timeline.py [log2n] [key_bits]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import particlerobotsimulations_b200 as prs
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 23
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n = 1 << log2n
L = prs.lib()
L.prs_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))
nt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
L.prs_sort_set_threads(nt)
plan = (C.c_int * 6)()
npass = L.prs_sort_plan(bits, n, plan)          # digits of this sort: 8 bits, or 9 where that saves a pass (prs_onesweep.cuh)
g = torch.Generator(device="cuda").manual_seed(1)
# cell-key-like input: lattice order, keys mostly ascending with local disorder
base = (torch.arange(n, device="cuda", dtype=torch.int64) * (1 << bits) // n)
keys = ((base + torch.randint(0, 1 << (bits // 2), (n,), device="cuda", generator=g)) % (1 << bits)).to(torch.int32)
vals = torch.arange(n, device="cuda", dtype=torch.int32)
ok, ov = torch.empty_like(keys), torch.empty_like(vals)
L.prs_sort_pairs(keys.data_ptr(), vals.data_ptr(), ok.data_ptr(), ov.data_ptr(), n, bits)
torch.cuda.synchronize()
tile = L.prs_sort_tile_size()                   # pairs per tile of that sort
tiles = (n + tile - 1) // tile
print(f"n=2^{log2n}, {bits}-bit keys: digits {[plan[i] for i in range(npass)]}, {tile} pairs per tile, {tiles} tiles")
tl = torch.zeros(npass * tiles * 8, dtype=torch.int64, device="cuda")
for it in range(3):
    L.prs_sort_set_timeline(C.c_void_p(tl.data_ptr()) if it == 2 else None)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); L.prs_sort_pairs(keys.data_ptr(), vals.data_ptr(), ok.data_ptr(), ov.data_ptr(), n, bits); b.record()
    torch.cuda.synchronize()
    print(f"run {it}: {a.elapsed_time(b)*1e3:.1f} us for {npass} passes + histogram, n=2^{log2n}, {nt} threads/tile")
L.prs_sort_set_timeline(None)
ref = torch.sort(keys.to(torch.int64) & 0xffffffff, stable=True)
assert torch.equal(ok.to(torch.int64) & 0xffffffff, ref.values) and torch.equal(ov.to(torch.int64), ref.indices), "sort mismatch"
t = tl.cpu().numpy().reshape(npass, tiles, 8).astype(np.float64)
names = ["keys+early counts", "prefix+scan+AGG", "lb request+rank->smem", "lb consume+publish", "barrier", "write out"]
for p in range(npass):
    x = t[p]
    t0 = x[:, 0].min()
    print(f"pass {p}: span {(x[:, 6].max() - t0)/1e3:.1f} us, tile life mean {(x[:, 6]-x[:, 0]).mean()/1e3:.2f} us; phases (mean us): " +
          ", ".join(f"{nm} {((x[:, i+1]-x[:, i]).mean())/1e3:.2f}" for i, nm in enumerate(names)))
    starts = np.sort(x[:, 0] - t0) / 1e3
    print("   tile start times (us) deciles:", np.round(np.percentile(starts, [0, 10, 25, 50, 75, 90, 100]), 1))
