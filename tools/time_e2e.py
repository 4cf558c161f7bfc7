import sys, os, importlib, time
sys.path.insert(0, '.')
import numpy as np
pkg = importlib.import_module('parallel-genomeseq_b200')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 151552
ref = pkg.synth.c3_reference(1_000_000)
ref_u8 = np.frombuffer(ref.encode('ascii'), dtype=np.uint8)
reads = pkg.synth.mutated_reads_fast(ref_u8, n, 150, seed=1, sub=0.01)
blob = np.ascontiguousarray(reads.reshape(-1)); offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(150))
e = pkg.Engine(0); e.set_scoring_match(0, 3, -3, 2); e.set_reference(ref)
for it in range(3):
    t0 = time.perf_counter(); e.stage((blob, offs), consensus=True, cons_stride=364); t1 = time.perf_counter()
    us = e.run(); t2 = time.perf_counter(); r = e.fetch(); t3 = time.perf_counter()
    print(f"stage {t1-t0:.3f}s run {t2-t1:.3f}s (device {us/1e6:.3f}s) fetch {t3-t2:.3f}s")
    t0 = time.perf_counter(); o = e.align((blob, offs), consensus=True, cons_stride=364, decode=False); t1 = time.perf_counter()
    print(f"align one-call {t1-t0:.3f}s device {o['device_us']/1e6:.3f}s")
import torch
torch.cuda.set_device(0)
blob_t = torch.from_numpy(blob.copy()).pin_memory(); blob_p = blob_t.numpy()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for it in range(3):
    t0 = time.perf_counter(); o = e.align((blob_p, offs), consensus=True, cons_stride=364, decode=False); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"[torch+pinned] align one-call {t1-t0:.3f}s device {o['device_us']/1e6:.3f}s")
