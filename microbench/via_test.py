import sys, numpy as np, time
sys.path.insert(0,'/root/repo')
import kmers_anno_b200 as ka
from kmers_anno_b200 import synth
from kmers_anno_b200.engine import pinned_array
fam=synth.Families(3000); kmers,roles=fam.table(3_000_000,K=8)
res,off,_=fam.batch(0,100,n_prot=4500,alloc=pinned_array)
with ka.Engine([0]) as eng:
    eng.db_load(kmers,roles,8)
    base=eng.annotate(res,off,5)
    codes,off32=eng.pack(res,off,alloc=pinned_array)
    for via in (-1,1):
        eng.set_option("ingest_via",via)
        for name,call in (("bytes",lambda: eng.annotate(res,off,5)),("packed",lambda: eng.annotate_packed(codes,off32,5))):
            call(); t=time.perf_counter(); got=call(); dt=(time.perf_counter()-t)*1e3
            print("via",via,name,"ms %.2f"%dt,"same",all(np.array_equal(x,y) for x,y in zip(got,base)), flush=True)
