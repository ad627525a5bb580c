import sys, time, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_pkg()
import torch
for wl, pairs in (("config1", 1_000_000), ("config2s", 1_000_000)):
    t=time.time()
    if wl=="config1":
        gb,go = pkg.synth.random_genomes(50, 3_000_000, seed=1)
    else:
        gb,go = pkg.synth.tree_genomes(100, 2_000_000, seed=1)
    rb,ro,_ = pkg.synth.paired_reads(gb,go,pairs,seed=2)
    print(wl,"gen",time.time()-t, flush=True)
    for cig in (False, True):
        al = pkg.Aligner(report_cigar=cig); al.set_debug_taps(False)
        t=time.time(); al.load_genomes(gb,go); print("load",time.time()-t)
        al.upload_reads(rb,ro)
        for i in range(3):
            t=time.time(); n=al.align_resident(); npairs=al.pair_batch(fetch=False); dt=time.time()-t
        tm=al.timings()
        print(wl, "cigar",cig, "step s",dt, "pairs/min M", pairs/dt*60/1e6)
        print(json.dumps({k:(round(v,3) if isinstance(v,float) else v) for k,v in tm.items()}))
        t=time.time(); res=al.align_batch(rb,ro,copy=False); pr=al.pair_batch(fetch=True,copy=False); print("e2e s",time.time()-t, "pack ms", al.timings()["ms_pack"], "d2h", al.timings()["ms_d2h"])
        al.close()
